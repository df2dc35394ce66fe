#!/bin/bash
# oracle/build_fixed.sh -- TEST INFRASTRUCTURE ONLY.  Builds the reference WITH the upstreamable fix of oracle/ub_fixes.py
#  the way a user would build it -- README one-liner, stock allocator, no special flags -- into
# oracle/_ref/nhw-enc-fixed and nhw-dec-fixed.  The patched copies live in a scratch directory and are deleted.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${NHW_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
[ -d "$REF/encoder" ] || { echo "build_fixed.sh: reference not found at $REF" >&2; exit 0; }
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT"
python3 "$HERE/ub_fixes.py" "$REF" "$TMP"
(cd "$TMP/encoder" && ${CC:-gcc} -O3 -w -ffp-contract=off *.c -o "$OUT/nhw-enc-fixed" -lm)
(cd "$TMP/decoder" && ${CC:-gcc} -O3 -w -ffp-contract=off *.c -o "$OUT/nhw-dec-fixed" -lm)
cp "$TMP/ub_fixes.patch" "$OUT/ub_fixes.patch"
echo "oracle/_ref: nhw-enc-fixed nhw-dec-fixed ub_fixes.patch"

#!/bin/bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Compiles the UNMODIFIED reference (default /root/reference) from the sources where
# they lie into oracle/_ref/ (git-ignored build output, travels to the GPU box):
#   libnhwref_enc.so / libnhwref_dec.so : canonical (zero-guard allocator) encoder and
#       decoder + in-memory glue + stage taps.  This is the parity oracle.
#   nhw-enc-canon / nhw-dec-canon       : the reference CLIs, canonical allocator.
#   nhw-enc-stock / nhw-dec-stock       : README one-liner build (gcc *.c -O3), stock malloc.
# The reference's own build system (CMake) is not used.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${NHW_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
TMP="$OUT/tmp"
if [ ! -d "$REF/encoder" ]; then
	echo "build_ref.sh: reference not found at $REF (prebuilt oracle/_ref is used as is)" >&2
	exit 0
fi
mkdir -p "$OUT" "$TMP"
CC="${CC:-gcc}"
# -ffp-contract=off: the reference's x86-64 gcc -O3 build has no FMA; keep it that way.
CFLAGS="-O3 -fPIC -w -ffp-contract=off"
# The canonical build also defines never-written AUTOMATIC variables as 0 (gcc >= 12: -ftrivial-auto-var-init=zero), exactly
# as the zero-guard allocator does for the heap: wavlts2packet reads one entry past its codebook list out of a local array
# (encoder/compress_pixel.c:58,412,446), and what the stack holds there -- leftovers of earlier frames, saved pointers whose
# bytes depend on the heap layout -- otherwise decides the last byte of tree1 on about one image in a few thousand (and
# differs between two calls on the same pixels from different threads).  A compiler flag, not a source change.
CANON="-ftrivial-auto-var-init=zero"
if ! echo 'int main(void){return 0;}' | $CC $CANON -x c - -o /dev/null 2>/dev/null; then
	echo "build_ref.sh: this gcc has no -ftrivial-auto-var-init; the canonical oracle keeps stack-dependent reads" >&2
	CANON=""
fi
WRAP="-Wl,--wrap=malloc,--wrap=calloc,--wrap=free,--wrap=exit"

ENC_SRCS="colorspace.c compress_pixel.c filters.c image_processing.c wavelet_filterbank.c"
DEC_SRCS="compress_pixel.c filters.c wavelet_filterbank.c"

# ---- encoder library (tap-instrumented nhw_encoder.c) ----
python3 "$HERE/make_tapped.py" "$REF/encoder/nhw_encoder.c" "$HERE/taps_enc.txt" "$TMP/enc_tapped.c"
objs=""
cobjs=""
for s in $ENC_SRCS; do
	$CC $CFLAGS -I"$REF/encoder" -c "$REF/encoder/$s" -o "$TMP/enc_${s%.c}.o"                 # stock (timing build)
	$CC $CFLAGS $CANON -I"$REF/encoder" -c "$REF/encoder/$s" -o "$TMP/encc_${s%.c}.o"         # canonical
	objs="$objs $TMP/enc_${s%.c}.o"
	cobjs="$cobjs $TMP/encc_${s%.c}.o"
done
# the q22/q23 side channel is computed inside wavelet_filterbank.c: the canonical library gets a tapped copy of it too
python3 "$HERE/make_tapped.py" "$REF/encoder/wavelet_filterbank.c" "$HERE/taps_enc.txt" "$TMP/enc_wfb_tapped.c"
$CC $CFLAGS $CANON -I"$REF/encoder" -I"$HERE" -c "$TMP/enc_wfb_tapped.c" -o "$TMP/enc_wavelet_filterbank_tapped.o"
tobjs="${cobjs/$TMP\/encc_wavelet_filterbank.o/$TMP/enc_wavelet_filterbank_tapped.o}"
$CC $CFLAGS $CANON -I"$REF/encoder" -I"$HERE" -c "$TMP/enc_tapped.c" -o "$TMP/enc_nhw_encoder.o"
$CC $CFLAGS -I"$REF/encoder" -I"$HERE" -c "$HERE/ref_enc_glue.c" -o "$TMP/enc_glue.o"
$CC $CFLAGS -c "$HERE/zguard.c" -o "$TMP/zguard.o"
$CC -shared -o "$OUT/libnhwref_enc.so" $tobjs "$TMP/enc_nhw_encoder.o" "$TMP/enc_glue.o" "$TMP/zguard.o" \
	$WRAP -Wl,-Bsymbolic -lm -lpthread
# timing build: the same reference objects with blocks padded but NOT zero-filled (zguard.c, NHW_NO_ZGUARD), so the
# CPU baseline is not slowed down by the canonicalising allocator and cannot fault on the reference's out-of-bounds
# reads.  Never used for parity.
$CC $CFLAGS -I"$REF/encoder" -c "$REF/encoder/nhw_encoder.c" -o "$TMP/enc_nhw_encoder_plain.o"
$CC $CFLAGS -DNHW_NO_ZGUARD -c "$HERE/zguard.c" -o "$TMP/zguard_exit_only.o"
$CC $CFLAGS -I"$REF/encoder" -I"$HERE" -DNHW_NO_TAPS -c "$HERE/ref_enc_glue.c" -o "$TMP/enc_glue_plain.o"
$CC -shared -o "$OUT/libnhwref_enc_stock.so" $objs "$TMP/enc_nhw_encoder_plain.o" "$TMP/enc_glue_plain.o" "$TMP/zguard_exit_only.o" \
	$WRAP -Wl,-Bsymbolic -lm -lpthread

# ---- decoder library ----
if [ -f "$HERE/taps_dec.txt" ]; then
	python3 "$HERE/make_tapped.py" "$REF/decoder/nhw_decoder.c" "$HERE/taps_dec.txt" "$TMP/dec_tapped.c"
	DEC_MAIN="$TMP/dec_tapped.c"
else
	DEC_MAIN="$REF/decoder/nhw_decoder.c"
fi
objs=""
for s in $DEC_SRCS; do
	$CC $CFLAGS $CANON -I"$REF/decoder" -c "$REF/decoder/$s" -o "$TMP/dec_${s%.c}.o"
	objs="$objs $TMP/dec_${s%.c}.o"
done
$CC $CFLAGS $CANON -I"$REF/decoder" -I"$HERE" -c "$DEC_MAIN" -o "$TMP/dec_nhw_decoder.o"
$CC $CFLAGS $CANON -I"$REF/decoder" -Dmain=nhwref_dec_cli_main -c "$REF/decoder/nhw_decoder_cli.c" -o "$TMP/dec_cli.o"
$CC $CFLAGS -I"$REF/decoder" -I"$HERE" -c "$HERE/ref_dec_glue.c" -o "$TMP/dec_glue.o"
$CC -shared -o "$OUT/libnhwref_dec.so" $objs "$TMP/dec_nhw_decoder.o" "$TMP/dec_cli.o" "$TMP/dec_glue.o" "$TMP/zguard.o" \
	$WRAP -Wl,-Bsymbolic -lm -lpthread

# ---- CLIs ----
(cd "$REF/encoder" && $CC -O3 -w -ffp-contract=off $CANON *.c "$TMP/zguard.o" -o "$OUT/nhw-enc-canon" $WRAP -lm)
(cd "$REF/decoder" && $CC -O3 -w -ffp-contract=off $CANON *.c "$TMP/zguard.o" -o "$OUT/nhw-dec-canon" $WRAP -lm)
(cd "$REF/encoder" && $CC -O3 -w *.c -o "$OUT/nhw-enc-stock" -lm)
(cd "$REF/decoder" && $CC -O3 -w *.c -o "$OUT/nhw-dec-stock" -lm)

rm -rf "$TMP"
echo "oracle/_ref built: $(ls "$OUT" | tr '\n' ' ')"

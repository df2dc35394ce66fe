#!/usr/bin/env python3
"""oracle/ub_fixes.py -- SURVEY.md section 8(f).2: the upstreamable fix for the reference's undefined reads, as a generator.

    python oracle/ub_fixes.py <reference root> <out dir>

writes patched COPIES of encoder/*.c and decoder/*.c into <out dir>/encoder and <out dir>/decoder (headers are copied as they
are), the new header nhw_defined_memory.h next to them, and <out dir>/ub_fixes.patch (a zero-context unified diff against
the reference tree, ready for `patch -p1`).  Reference sources are never committed to this repository: the copies and the
patch are build intermediates (oracle/build_fixed.sh compiles them and tests/test_oracle_cpu.py compares the result with the
canonical oracle).

What the patch does -- and all it does:

  1. nhw_defined_memory.h: malloc / calloc / free of the codec go through three small functions that hand out zero-filled
     blocks with a zeroed margin on both sides.  Every out-of-bounds read listed in SURVEY.md Appendix C stays within a few
     rows of its block, and every read of never-written heap memory (tree1[16384..], nhw_kernel borders ...) sees zeros:
     the output no longer depends on the allocator, on earlier allocations or on the heap layout.
  2. every .c file of the two programs includes that header after its own includes.
  3. wavlts2packet's local codebook array is zero-initialised (encoder/compress_pixel.c:58): its run-length loop looks one
     entry past the list it has just de-interleaved (:412, :446).

The result, built with the STOCK allocator and no special flags, is byte-identical to the canonical oracle
(zero-guard allocator at link time + -ftrivial-auto-var-init=zero) on the inputs the test runs; that is the claim
"these reads are the only reason the stock output varies".  It is a minimal, behaviour-defining patch; bounds checks at each
of the ~25 sites would be the more surgical follow-up.
"""
import difflib
import os
import re
import shutil
import sys

HEADER = r'''/* nhw_defined_memory.h -- gives the codec's out-of-bounds and uninitialised heap reads a defined value (0).
 * Every block is zero-filled and carries NHW_DM_MARGIN zero bytes on both sides. */
#ifndef NHW_DEFINED_MEMORY_H
#define NHW_DEFINED_MEMORY_H
#include <stdlib.h>
#define NHW_DM_MARGIN 65536
static inline void *nhw_dm_alloc(size_t n)
{
	char *p = (char *)calloc(1, n + 2 * (size_t)NHW_DM_MARGIN);
	return p ? p + NHW_DM_MARGIN : NULL;
}
static inline void *nhw_dm_calloc(size_t a, size_t b) { return nhw_dm_alloc(a * b); }
static inline void nhw_dm_free(void *p)
{
	if (p) free((char *)p - NHW_DM_MARGIN);
}
#define malloc(n) nhw_dm_alloc(n)
#define calloc(a, b) nhw_dm_calloc(a, b)
#define free(p) nhw_dm_free(p)
#endif
'''


def patch_source(text, rel):
    lines = text.split("\n")
    last_inc = max((i for i, l in enumerate(lines) if l.lstrip().startswith("#include")), default=-1)
    lines.insert(last_inc + 1, '#include "nhw_defined_memory.h"')
    out = "\n".join(lines)
    if rel == "encoder/compress_pixel.c":
        out, n = re.subn(r"\bcodebook\[580\]", "codebook[580]={0}", out, count=1)
        assert n == 1, "codebook declaration not found"
    return out


def main():
    ref, dst = sys.argv[1:3]
    patch = []
    for sub in ("encoder", "decoder"):
        os.makedirs(os.path.join(dst, sub), exist_ok=True)
        with open(os.path.join(dst, sub, "nhw_defined_memory.h"), "w") as f:
            f.write(HEADER)
        patch += list(difflib.unified_diff([], HEADER.splitlines(), "/dev/null", "b/%s/nhw_defined_memory.h" % sub, n=0, lineterm=""))
        for name in sorted(os.listdir(os.path.join(ref, sub))):
            src = os.path.join(ref, sub, name)
            if name.endswith(".h"):
                shutil.copy(src, os.path.join(dst, sub, name))
            elif name.endswith(".c"):
                old = open(src, encoding="latin-1").read()
                new = patch_source(old, sub + "/" + name)
                with open(os.path.join(dst, sub, name), "w", encoding="latin-1") as f:
                    f.write(new)
                patch += list(difflib.unified_diff(old.split("\n"), new.split("\n"), "a/%s/%s" % (sub, name), "b/%s/%s" % (sub, name),
                                                   n=0, lineterm=""))
    with open(os.path.join(dst, "ub_fixes.patch"), "w", encoding="latin-1") as f:
        f.write("\n".join(patch) + "\n")
    print("ub_fixes: patched copies and ub_fixes.patch (%d lines) written to %s" % (len(patch), dst))


if __name__ == "__main__":
    main()

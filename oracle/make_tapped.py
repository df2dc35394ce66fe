#!/usr/bin/env python3
"""oracle/make_tapped.py -- TEST INFRASTRUCTURE ONLY.

Writes a scratch copy of one reference source file with NHW_TAP(...) dump statements
inserted before the line numbers listed in a tap spec (oracle/taps_*.txt).  The copy is
a build intermediate: it is written under oracle/_ref/tmp/, compiled, and deleted by
oracle/build_ref.sh -- reference sources are never committed to this repository.

usage: make_tapped.py <src.c> <spec.txt> <out.c>
"""
import os
import sys


def main():
    src, spec, out = sys.argv[1:4]
    base = os.path.basename(src)
    inserts = {}
    for raw in open(spec):
        raw = raw.strip()
        if not raw or raw.startswith("#"):
            continue
        fname, line, code = raw.split(":", 2)
        if fname.strip() != base:
            continue
        inserts.setdefault(int(line), []).append(code.strip())
    lines = open(src, encoding="latin-1").read().split("\n")
    res = ['#include "tap.h"', '#line 1 "%s"' % src]
    for no, text in enumerate(lines, start=1):
        for code in inserts.get(no, []):
            res.append("{ %s }" % code)
            res.append('#line %d "%s"' % (no, src))
        res.append(text)
    with open(out, "w", encoding="latin-1") as f:
        f.write("\n".join(res))


if __name__ == "__main__":
    main()

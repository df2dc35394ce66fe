#!/usr/bin/env python
"""tests/golden/make_golden.py -- regenerates tests/golden/golden.json and the .nhw fixtures from the
canonical build of the reference (oracle/_ref, built by oracle/build_ref.sh from /root/reference).

    python tests/golden/make_golden.py

For each (generator, seed, quality) case: length + md5 of the .nhw stream the reference encoder writes
and md5 of the 786432 pixel bytes the reference decoder writes for that stream.  Two complete streams
are committed as files so the container parser and the decoder can be exercised without the oracle.
The reference ships no test vectors of its own; these are outputs of the
reference itself, run in the build container."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from nhwcodec_b200 import synth  # noqa: E402
from oracle import refbind  # noqa: E402
from test_oracle_cpu import smooth_pixels  # noqa: E402

CASES = [("smooth", 0), ("natural", 1000), ("natural", 4242), ("textured", 1002), ("noise", 5)]
QUALITIES = [17, 18, 19, 20, 21, 22, 23]
FILES = [("smooth", 0, 20), ("natural", 1000, 23)]


def pixels(kind, seed):
    if kind == "smooth":
        return smooth_pixels()
    return {"natural": synth.natural, "textured": synth.textured, "noise": synth.noise}[kind](seed)


def main():
    out = {"note": "canonical (zero-guard) reference build, commit 582927d; see make_golden.py", "cases": []}
    for kind, seed in CASES:
        pix = pixels(kind, seed)
        for q in QUALITIES:
            s = refbind.ref_encode(pix, q)
            d = refbind.ref_decode(s)
            out["cases"].append({"kind": kind, "seed": seed, "q": q, "pixels_md5": hashlib.md5(pix.tobytes()).hexdigest(),
                                 "nhw_len": len(s), "nhw_md5": hashlib.md5(s).hexdigest(),
                                 "decoded_md5": hashlib.md5(d.tobytes()).hexdigest()})
            if (kind, seed, q) in FILES:
                with open(os.path.join(HERE, "%s_%d_q%d.nhw" % (kind, seed, q)), "wb") as f:
                    f.write(s)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("%d cases" % len(out["cases"]))


if __name__ == "__main__":
    main()

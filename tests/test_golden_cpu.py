"""The committed golden vectors (tests/golden/, produced by make_golden.py from the canonical build of
the reference) against the oracle as built on this machine, and the host-side container parser against
the committed .nhw files.  No GPU needed."""
import hashlib
import json
import os

import numpy as np
import pytest

from nhwcodec_b200 import synth
from test_oracle_cpu import smooth_pixels

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))["cases"]


def pixels(kind, seed):
    if kind == "smooth":
        return smooth_pixels()
    return {"natural": synth.natural, "textured": synth.textured, "noise": synth.noise}[kind](seed)


def test_generators_match_golden_inputs():
    seen = set()
    for c in GOLD:
        key = (c["kind"], c["seed"])
        if key in seen:
            continue
        seen.add(key)
        assert hashlib.md5(pixels(*key).tobytes()).hexdigest() == c["pixels_md5"], key


@pytest.mark.parametrize("q", [17, 20, 23])
def test_oracle_reproduces_golden(ref, q):
    for c in GOLD:
        if c["q"] != q:
            continue
        s = ref.ref_encode(pixels(c["kind"], c["seed"]), q)
        assert (len(s), hashlib.md5(s).hexdigest()) == (c["nhw_len"], c["nhw_md5"]), c
        assert hashlib.md5(ref.ref_decode(s).tobytes()).hexdigest() == c["decoded_md5"], c


def test_committed_streams_match_golden():
    for name in ("smooth_0_q20.nhw", "natural_1000_q23.nhw"):
        kind, seed, q = name[:-4].split("_")
        data = open(os.path.join(HERE, "golden", name), "rb").read()
        c = next(x for x in GOLD if (x["kind"], str(x["seed"]), "q%d" % x["q"]) == (kind, seed, q))
        assert (len(data), hashlib.md5(data).hexdigest()) == (c["nhw_len"], c["nhw_md5"])
        assert data[1] == c["q"]          # quality byte

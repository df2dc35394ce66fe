"""GPU encode / decode against the committed golden vectors (tests/golden/golden.json): no oracle involved,
so these also run where oracle/_ref is absent."""
import hashlib
import json
import os

import numpy as np
import pytest

from test_golden_cpu import GOLD, pixels

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("q", [17, 18, 19, 20, 21, 22, 23])
def test_encode_matches_golden(codec, q):
    cases = [c for c in GOLD if c["q"] == q]
    imgs = np.stack([pixels(c["kind"], c["seed"]) for c in cases])
    streams, status = codec.encode(imgs, q)
    assert (status == 0).all(), status
    for c, s in zip(cases, streams):
        assert (len(s), hashlib.md5(s).hexdigest()) == (c["nhw_len"], c["nhw_md5"]), c


def test_decode_committed_streams(codec):
    names = ("smooth_0_q20.nhw", "natural_1000_q23.nhw")
    streams = [open(os.path.join(HERE, "golden", n), "rb").read() for n in names]
    rgb, status = codec.decode(streams)
    assert (status == 0).all()
    for n, px in zip(names, rgb):
        kind, seed, q = n[:-4].split("_")
        c = next(x for x in GOLD if (x["kind"], str(x["seed"]), "q%d" % x["q"]) == (kind, seed, q))
        assert hashlib.md5(px.tobytes()).hexdigest() == c["decoded_md5"], n


def test_round_trip_properties(codec):
    """size-independent properties on a larger batch: decode(encode(x)) is deterministic, identical images give
    identical streams wherever they sit in the batch, and the decoded picture is close to the input (PSNR)"""
    from nhwcodec_b200 import synth
    base = np.stack([synth.natural(6000 + i) for i in range(8)])
    imgs = np.concatenate([base, base[::-1], base])          # 24 images, repeated content at different slots
    s, st = codec.encode(imgs, 20)
    assert (st == 0).all()
    for i in range(8):
        assert s[i] == s[15 - i] == s[16 + i]
    rgb, st = codec.decode(s)
    assert (st == 0).all()
    rgb2, _ = codec.decode(s)
    assert np.array_equal(rgb, rgb2)
    for i in range(8):
        err = rgb[i].astype(np.float64) - imgs[i].astype(np.float64)
        psnr = 10 * np.log10(255.0 ** 2 / max(np.mean(err * err), 1e-9))
        assert psnr > 30.0, (i, psnr)

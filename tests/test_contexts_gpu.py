"""Contexts are independent: several in one process, on the same GPU and (where the box has more than one) on
different GPUs.  The > 48 KB shared-memory opt-ins and the constant tables are per-device state set up by every
nhw_create (they used to sit behind process-wide flags, so a second device never got them)."""
import numpy as np
import pytest

from nhwcodec_b200 import synth

pytestmark = pytest.mark.gpu


def _imgs(n, seed0):
    fs = [synth.natural, synth.textured, synth.noise]
    return np.stack([fs[i % 3](seed0 + i) for i in range(n)])


def test_two_contexts_one_process(ref):
    import torch
    from nhwcodec_b200 import Codec
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    devices = [0, 0] + ([1] if torch.cuda.device_count() > 1 else [])
    codecs = [Codec(device=d, max_batch=4) for d in devices]
    try:
        imgs = _imgs(6, 8800)
        want = {q: [ref.ref_encode(imgs[i], q) for i in range(6)] for q in (20, 23, 9)}
        for q in (20, 23, 9):
            for c in codecs:   # interleaved use of the contexts
                streams, status = c.encode(imgs, q)
                assert (status == 0).all()
                assert streams == want[q], (q, c)
        for c in codecs:
            rgb, st = c.decode(want[20])
            assert (st == 0).all()
            assert np.array_equal(rgb[0], ref.ref_decode(want[20][0]))
    finally:
        for c in codecs:
            c.close()


def test_contexts_from_threads(ref):
    """two host threads, each with its own context, encoding at the same time"""
    import threading
    import torch
    from nhwcodec_b200 import Codec
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    imgs = _imgs(8, 8900)
    want = [ref.ref_encode(imgs[i], 20) for i in range(8)]
    out = [None, None]

    def work(k):
        c = Codec(device=0, max_batch=4)
        try:
            for _ in range(3):
                streams, status = c.encode(imgs, 20)
                assert (status == 0).all()
            out[k] = streams
        finally:
            c.close()

    ts = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert out[0] == want and out[1] == want


def test_concurrent_creates_leave_the_device_tables_intact(ref):
    """contexts created from several threads at once (each nhw_create uploads the per-device tables) while an older
    context keeps decoding: the decoder's prefix-code table used to be rebuilt in a shared host array by every create, so
    two creates at once could upload a half-zeroed table and break every context's decode of the rarer codes"""
    import threading
    import torch
    from nhwcodec_b200 import Codec
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    old = Codec(device=0, max_batch=2)
    noise = np.stack([synth.noise(9103), synth.noise(9106)])
    streams = [ref.ref_encode(noise[i], 1) for i in range(2)]      # big alphabets: long, rare codes
    want = [ref.ref_decode(s) for s in streams]
    stop = [False]

    def churn():
        while not stop[0]:
            Codec(device=0, max_batch=1).close()

    ts = [threading.Thread(target=churn) for _ in range(3)]
    for t in ts:
        t.start()
    try:
        for _ in range(12):
            rgb, st = old.decode(streams)
            assert (st == 0).all(), st
            assert np.array_equal(rgb[0], want[0]) and np.array_equal(rgb[1], want[1])
    finally:
        stop[0] = True
        for t in ts:
            t.join()
        old.close()

"""GPU parity of the full encode: .nhw bytes from libnhw_cuda must be identical to the
canonical build of the reference encoder (oracle/_ref) on the same pixels."""
import numpy as np
import pytest

from nhwcodec_b200 import synth

pytestmark = pytest.mark.gpu


def _mixed(n, seed0):
    fs = [synth.natural, synth.textured, synth.noise]
    return np.stack([fs[i % 3](seed0 + i) for i in range(n)])


@pytest.mark.parametrize("q", list(range(1, 24)))
def test_encode_bit_exact(codec, ref, q):
    imgs = _mixed(6, 7000 + q)
    streams, status = codec.encode(imgs, q)
    assert (status == 0).all(), status
    for i in range(imgs.shape[0]):
        want = ref.ref_encode(imgs[i], q)
        got = streams[i]
        assert len(got) == len(want), (q, i, len(got), len(want))
        if got != want:
            a = np.frombuffer(got, np.uint8)
            b = np.frombuffer(want, np.uint8)
            d = np.flatnonzero(a != b)
            raise AssertionError("q=%d image %d: %d bytes differ, first at %d" % (q, i, d.size, d[0]))


def test_encode_chunking_and_device_api(codec, ref):
    """more images than max_batch (16): chunked host path == device-resident path == oracle"""
    import torch
    imgs = _mixed(40, 8100)
    streams, status = codec.encode(imgs, 20)
    assert (status == 0).all()
    t = torch.from_numpy(imgs).cuda()
    out = torch.zeros((40, 1 << 19), dtype=torch.uint8, device="cuda")
    ln = torch.zeros(40, dtype=torch.int32, device="cuda")
    st = torch.zeros(40, dtype=torch.int32, device="cuda")
    codec.encode_device(t, 20, out, ln, st)
    ln = ln.cpu().numpy()
    outc = out.cpu().numpy()
    for i in range(40):
        assert streams[i] == outc[i, :ln[i]].tobytes(), i
    for i in (0, 1, 2, 17, 39):
        assert streams[i] == ref.ref_encode(imgs[i], 20), i


@pytest.mark.parametrize("q", [1, 8, 12, 16, 17, 20, 21, 22, 23])
def test_smooth_known_answer(codec, q):
    """size and md5 of the canonical .nhw of the formula-defined image"""
    import hashlib
    from test_oracle_cpu import KAT, smooth_pixels
    streams, status = codec.encode(smooth_pixels()[None, :], q)
    assert status[0] == 0
    assert len(streams[0]) == KAT[q][0]
    assert hashlib.md5(streams[0]).hexdigest() == KAT[q][1]


@pytest.mark.parametrize("q", [0, 24, -1])
def test_unsupported_quality_is_an_error(codec, q):
    """-q0 is accepted by the reference CLI but leaves its tables undefined: refused here"""
    from nhwcodec_b200 import NhwError
    with pytest.raises(NhwError):
        codec.encode(_mixed(1, 1), q)


def test_low_quality_device_api_matches_host_api(codec, ref):
    """q <= 16 through the device-resident entry point (single sub-chunk + chroma side stream) and the chunked host one"""
    import torch
    imgs = _mixed(20, 8300)
    for q in (3, 10, 14):
        streams, status = codec.encode(imgs, q)
        assert (status == 0).all()
        t = torch.from_numpy(imgs).cuda()
        out = torch.zeros((20, 1 << 19), dtype=torch.uint8, device="cuda")
        ln = torch.zeros(20, dtype=torch.int32, device="cuda")
        st = torch.zeros(20, dtype=torch.int32, device="cuda")
        codec.encode_device(t, q, out, ln, st)
        ln = ln.cpu().numpy()
        outc = out.cpu().numpy()
        for i in range(20):
            assert streams[i] == outc[i, :ln[i]].tobytes(), (q, i)
        for i in (0, 7, 19):
            assert streams[i] == ref.ref_encode(imgs[i], q), (q, i)

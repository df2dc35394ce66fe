"""GPU parity of the encoder front end (colour + 4:2:0, pre-sharpening, 2-level DWT) against
stage taps of the compiled canonical reference.  Bit-exact: integer planes must be identical."""
import numpy as np
import pytest

from nhwcodec_b200 import synth

pytestmark = pytest.mark.gpu


def _imgs():
    return np.stack([synth.natural(1000), synth.natural(1001), synth.textured(1002), synth.noise(5)])


def test_synth_device_matches_host(codec):
    import torch
    for kind, f in ((0, synth.natural), (1, synth.noise), (2, synth.textured)):
        t = torch.empty((2, 786432), dtype=torch.uint8, device="cuda")
        codec.synth(t, 4242, kind)
        got = t.cpu().numpy()
        for i in range(2):
            assert np.array_equal(got[i], f(4242 + i)), (kind, i)


@pytest.mark.parametrize("q", [20, 23, 19, 18, 17, 16, 9, 1])
def test_colorspace(codec, ref, q):
    """the stage kernel of front.cu (nhw_stage_colorspace_device): the product path at q <= 16, a stage-level check of the
    shared per-pixel arithmetic (color_core.cuh) at q >= 17, where the product kernel is the fused one (test_frontend_vs_taps)"""
    import torch
    imgs = _imgs()
    n = imgs.shape[0]
    t = torch.from_numpy(imgs).cuda()
    y = torch.empty((n, 512, 512), dtype=torch.int16, device="cuda")
    u = torch.empty((n, 256, 256), dtype=torch.uint8, device="cuda")
    v = torch.empty((n, 256, 256), dtype=torch.uint8, device="cuda")
    codec.stage_colorspace(t, q, False, y, u, v)
    for i in range(n):
        Y, U, V = ref.ref_colorspace(imgs[i], q)
        assert np.array_equal(y[i].cpu().numpy(), Y), (q, i, "Y")
        assert np.array_equal(u[i].cpu().numpy(), U), (q, i, "U")
        assert np.array_equal(v[i].cpu().numpy(), V), (q, i, "V")


def test_color_fast_path_exhaustive(codec):
    """the fused front end's integer q>=20 colour transform == the IEEE form (itself checked against
    the reference above) on every one of the 2^24 RGB triples"""
    assert codec.color_check() == 0


@pytest.mark.parametrize("q", [20, 17, 21, 16, 14, 10, 7, 5, 1])
def test_pre_processing(codec, ref, q):
    import torch
    imgs = _imgs()
    n = imgs.shape[0]
    t = torch.from_numpy(imgs).cuda()
    y = torch.empty((n, 512, 512), dtype=torch.int16, device="cuda")
    codec.stage_colorspace(t, q, True, y, None, None)
    for i in range(n):
        Y, _, _ = ref.ref_colorspace(imgs[i], q)
        want = ref.ref_pre_processing(Y, q)
        got = y[i].cpu().numpy()
        bad = np.argwhere(got != want)
        assert bad.size == 0, (q, i, bad[:5], got[tuple(bad[0])], want[tuple(bad[0])])


@pytest.mark.parametrize("q", [17, 18, 19, 20, 21, 22, 23, 16, 14, 13, 8, 2])
def test_frontend_vs_taps(codec, ref, q):
    """the PRODUCT front end (q >= 17: the fused k_front_luma in each of its colour modes, with and without pre-sharpening and
    the kept first pass; q <= 16: colour + state machine + plane-fed analysis) against the reference's stage taps"""
    import torch
    imgs = _imgs()
    n = imgs.shape[0]
    t = torch.from_numpy(imgs).cuda()
    yp = torch.empty((n, 512, 512), dtype=torch.int16, device="cuda")
    yl = torch.empty((n, 256, 256), dtype=torch.int16, device="cuda")
    cp = torch.empty((n, 2, 256, 256), dtype=torch.int16, device="cuda")
    cl = torch.empty((n, 2, 128, 128), dtype=torch.int16, device="cuda")
    codec.stage_frontend(t, q, yp, yl, cp, cl)
    for i in range(n):
        _, taps = ref.ref_encode_taps(imgs[i], q)
        want = taps["y_dwt2_proc"].view(np.int16).reshape(512, 512)
        got = yp[i].cpu().numpy()
        bad = np.argwhere(got != want)
        assert bad.size == 0, ("y_proc", q, i, bad[:5])
        assert np.array_equal(yl[i].cpu().numpy(), taps["y_ll1"].view(np.int16).reshape(256, 256)), ("ll1", q, i)
        for k, name in enumerate("uv"):
            want = taps[name + "_dwt2_proc"].view(np.int16).reshape(256, 256)
            got = cp[i, k].cpu().numpy()
            bad = np.argwhere(got != want)
            assert bad.size == 0, (name, q, i, bad[:5])
            assert np.array_equal(cl[i, k].cpu().numpy(), taps[name + "_ll1"].view(np.int16).reshape(128, 128))

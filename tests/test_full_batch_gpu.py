"""BASELINE.json configs[1] at full size (batch 4096, -q20) on the GPU, checked through properties that do not need
4096 oracle encodes: every image encodes (status 0), two runs give identical bytes, and a 1-in-128 sample equals the
compiled reference.  A 512-image slice then goes through the HOST API of a smaller context (16 sub-chunks of 32 images on
four streams) and must give the same bytes, goes back through the decoder and is compared with the reference decoder on
a sample.  test_host_api_waves_and_sub_chunks covers more images than max_batch (several waves of sub-chunks)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = 4096


def _digest(out, lens):
    import torch
    idx = torch.arange(out.shape[1], device=out.device)[None, :]
    masked = torch.where(idx < lens[:, None], out, torch.zeros_like(out)).to(torch.int64)
    w = (idx % 251 + 1).to(torch.int64)
    return (masked * w).sum(dim=1)          # per-image weighted checksum, on the device


def test_batch_4096_properties(ref):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from nhwcodec_b200 import Codec, synth
    big = Codec(device=0, max_batch=N)
    rgb = torch.empty((N, 786432), dtype=torch.uint8, device="cuda")
    big.synth(rgb, 1000, 0)
    out = torch.zeros((N, 1 << 19), dtype=torch.uint8, device="cuda")
    lens = torch.zeros(N, dtype=torch.int32, device="cuda")
    st = torch.zeros(N, dtype=torch.int32, device="cuda")
    big.encode_device(rgb, 20, out, lens, st)
    assert int((st != 0).sum()) == 0
    d1, l1 = _digest(out, lens), lens.clone()
    out.zero_()
    big.encode_device(rgb, 20, out, lens, st)
    assert torch.equal(l1, lens) and torch.equal(d1, _digest(out, lens)), "encode is not deterministic"
    lens_h = lens.cpu().numpy()
    for i in range(0, N, 128):              # 32 images against the compiled reference
        want = ref.ref_encode(synth.natural(1000 + i), 20)
        got = out[i, : int(lens_h[i])].cpu().numpy().tobytes()
        assert got == want, i
    # the host API of a smaller context: chunks of 1024, four lanes each
    small = Codec(device=0, max_batch=1024)
    sl = slice(1536, 1536 + 512)
    pix = rgb[sl].cpu().numpy()
    streams, status = small.encode(pix, 20)
    assert (status == 0).all()
    for k in range(0, 512, 37):
        assert streams[k] == out[1536 + k, : int(lens_h[1536 + k])].cpu().numpy().tobytes(), k
    back, dstat = small.decode(streams)
    assert (dstat == 0).all()
    for k in range(0, 512, 64):
        assert np.array_equal(back[k], ref.ref_decode(streams[k])), k
    err = back.astype(np.float32) - pix.astype(np.float32)
    psnr = 10 * np.log10(255.0 ** 2 / np.maximum((err * err).mean(axis=1), 1e-9))
    assert psnr.min() > 30.0, psnr.min()
    big.close()
    small.close()


def test_host_api_waves_and_sub_chunks(ref):
    """130 images through a context with max_batch 48: three waves (48 + 48 + 34), each cut into 16 (encode) / 8 (decode)
    sub-chunks of uneven size on four streams; host-API bytes == device-API bytes == oracle on a sample, in image order"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from nhwcodec_b200 import Codec, synth
    n = 130
    c = Codec(device=0, max_batch=48)
    fs = [synth.natural, synth.textured, synth.noise]
    imgs = np.stack([fs[i % 3](7700 + i) for i in range(n)])
    for q in (20, 9):
        streams, status = c.encode(imgs, q)
        assert (status == 0).all()
        t = torch.from_numpy(imgs).cuda()
        out = torch.zeros((n, 1 << 19), dtype=torch.uint8, device="cuda")
        ln = torch.zeros(n, dtype=torch.int32, device="cuda")
        st = torch.zeros(n, dtype=torch.int32, device="cuda")
        c.encode_device(t, q, out, ln, st)
        lh = ln.cpu().numpy()
        for i in range(n):
            assert streams[i] == out[i, : int(lh[i])].cpu().numpy().tobytes(), (q, i)
        for i in (0, 47, 48, 95, 96, 129):
            assert streams[i] == ref.ref_encode(imgs[i], q), (q, i)
        back, dstat = c.decode(streams)
        assert (dstat == 0).all()
        for i in (0, 5, 47, 48, 53, 96, 129):
            assert np.array_equal(back[i], ref.ref_decode(streams[i])), (q, i)
    c.close()

"""BASELINE.json configs[1] at full size (batch 4096, -q20) on the GPU, checked through properties that do not need
4096 oracle encodes: every image encodes (status 0), two runs give identical bytes, a chunked context
(max_batch 1024 -> 4 chunks, and 4 lanes per chunk) gives the same bytes as one 4096-image chunk, and a
1-in-128 sample equals the compiled reference.  Then a 512-image slice goes back through the decoder and
is compared with the reference decoder on a sample."""
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

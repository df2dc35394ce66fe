"""GPU parity of the decoder: pixels from libnhw_cuda's nhw_decode_batch must equal what the
reference decoder (oracle/_ref, nhw-dec's decode_image + write_image_bmp) writes for the same
.nhw stream.  Streams come from the canonical reference encoder and from our own encoder."""
import numpy as np
import pytest

from nhwcodec_b200 import synth

pytestmark = pytest.mark.gpu


def _mixed(n, seed0):
    fs = [synth.natural, synth.textured, synth.noise]
    return np.stack([fs[i % 3](seed0 + i) for i in range(n)])


@pytest.mark.parametrize("q", list(range(1, 24)))
def test_decode_bit_exact(codec, ref, q):
    imgs = _mixed(6, 9100 + q)
    streams = [ref.ref_encode(imgs[i], q) for i in range(imgs.shape[0])]
    rgb, status = codec.decode(streams)
    assert (status == 0).all(), status
    for i, s in enumerate(streams):
        want = ref.ref_decode(s)
        d = np.flatnonzero(rgb[i] != want)
        assert d.size == 0, "q=%d image %d: %d bytes differ, first at %d" % (q, i, d.size, d[0])


def test_round_trip_own_streams_and_chunking(codec, ref):
    """encode on the GPU, decode on the GPU (40 > max_batch 16: chunked), compare with the
    reference decoder on the same streams; mixed qualities in one decode batch."""
    imgs = _mixed(20, 9300)
    s20, st = codec.encode(imgs, 20)
    s18, st2 = codec.encode(imgs, 23)
    assert (st == 0).all() and (st2 == 0).all()
    streams = s20 + s18
    rgb, status = codec.decode(streams)
    assert (status == 0).all()
    for i in (0, 1, 2, 19, 20, 21, 39):
        assert np.array_equal(rgb[i], ref.ref_decode(streams[i])), i


@pytest.mark.parametrize("q", [1, 8, 12, 16, 17, 20, 22, 23])
def test_decode_known_answer(codec, ref, q):
    """md5 of the BMP nhw-dec writes for the formula-defined image"""
    import hashlib
    from test_oracle_cpu import KAT, smooth_pixels
    stream = ref.ref_encode(smooth_pixels(), q)
    rgb, status = codec.decode([stream])
    assert status[0] == 0
    hdr = bytes([66, 77, 54, 0, 12, 0, 0, 0, 0, 0, 54, 0, 0, 0, 40, 0, 0, 0, 0, 2, 0, 0, 0, 2, 0, 0, 1, 0, 24, 0,
                 0, 0, 0, 0, 0, 0, 12, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    assert hashlib.md5(hdr + rgb[0].tobytes()).hexdigest() == KAT[q][2]


def test_decode_rejects_garbage(codec):
    rgb, status = codec.decode([b"\x09" + b"\0" * 200, b"\0" * 10])
    assert (status != 0).all()


def test_decode_mixed_low_and_high_quality_batch(codec, ref):
    """streams of every quality in ONE decode batch (each carries its own quality byte)"""
    imgs = _mixed(23, 9500)
    streams = [ref.ref_encode(imgs[i], i + 1) for i in range(23)]
    rgb, status = codec.decode(streams)
    assert (status == 0).all(), status
    for i, s in enumerate(streams):
        assert np.array_equal(rgb[i], ref.ref_decode(s)), "q=%d" % (i + 1)


def test_decode_hostile_headers(codec, ref):
    """section lengths beyond the decode workspace, truncated streams, lying lengths: per-stream error status, the
    good stream next to them still decodes (dec_parse.h)"""
    import struct
    img = _mixed(1, 9600)[0]
    good = ref.ref_encode(img, 19)
    bad = []
    b = bytearray(good)
    # q19 header: byte0,q, tree1(2) tree2(2) data1(4) data2(4) tree_end(2) exw(2) res1_len(2) res3_len(2) res3_bit_len(2) ...
    struct.pack_into("<H", b, 22, 65535)          # res3_bit_len
    bad.append(bytes(b) + b"\0" * 300000)
    b = bytearray(good); struct.pack_into("<H", b, 14, 65535); bad.append(bytes(b))   # tree_end
    bad.append(good[:36])                                                               # truncated after the header
    bad.append(good[:20])                                                               # truncated inside the header
    b = bytearray(good); b[1] = 23; bad.append(bytes(b[:37]))                           # quality byte lies about the header size
    b = bytearray(good); b[1] = 0; bad.append(bytes(b))                                 # q0
    rgb, status = codec.decode(bad + [good])
    assert (status[:-1] != 0).all(), status
    assert status[-1] == 0 and np.array_equal(rgb[-1], ref.ref_decode(good))


def test_decode_corrupted_ll_section_is_contained(codec, ref):
    """random damage inside the LL byte section (and its side list) of otherwise valid streams: the parallel LL decoder
    (kd_ll_parallel) must stay inside its buffers whatever the codes say -- no CUDA error, the context keeps working, and
    an undamaged stream in the same batch still decodes exactly"""
    from nhwcodec_b200 import container
    rng = np.random.default_rng(5)
    img = _mixed(1, 9700)[0]
    batch = []
    for q in (20, 12, 23):
        good = ref.ref_encode(img, q)
        h, sec = container.parse_nhw(good)
        start = h["header_bytes"]
        for name, data in sec.items():
            if name == "ch_res":
                break
            start += len(data)
        n = len(sec["ch_res"])
        for k in range(20):
            b = bytearray(good)
            for _ in range(1 + k % 7):
                b[start + int(rng.integers(0, n))] = int(rng.integers(0, 256))
            if k % 5 == 0:        # also cut the section short by lying about where the words start is not possible; zero a run instead
                a = start + int(rng.integers(0, max(n - 64, 1)))
                b[a:a + 64] = bytes(64)
            batch.append(bytes(b))
        batch.append(good)
    rgb, status = codec.decode(batch)
    for i in (20, 41, 62):
        assert status[i] == 0 and np.array_equal(rgb[i], ref.ref_decode(batch[i])), i
    again, st2 = codec.decode([batch[20]])
    assert st2[0] == 0 and np.array_equal(again[0], rgb[20])


def test_decode_device_api(codec, ref):
    """device-resident decode (headers walked on the device): slots as written by the device encoder, and streams
    packed back to back; more streams than max_batch (16), mixed qualities, one corrupt stream in the middle"""
    import torch
    imgs = _mixed(40, 9700)
    t = torch.from_numpy(imgs).cuda()
    slots = torch.zeros((40, 1 << 19), dtype=torch.uint8, device="cuda")
    ln = torch.zeros(40, dtype=torch.int32, device="cuda")
    st = torch.zeros(40, dtype=torch.int32, device="cuda")
    codec.encode_device(t[:20], 20, slots[:20], ln[:20], st[:20])
    codec.encode_device(t[20:], 7, slots[20:], ln[20:], st[20:])
    assert int((st != 0).sum()) == 0
    slots[13, 1] = 0                      # quality byte 0: refused
    rgb = torch.empty((40, 786432), dtype=torch.uint8, device="cuda")
    dst = torch.zeros(40, dtype=torch.int32, device="cuda")
    codec.decode_device(slots, ln, rgb, dst)
    dst_h = dst.cpu().numpy()
    assert dst_h[13] != 0 and (np.delete(dst_h, 13) == 0).all(), dst_h
    lens = ln.cpu().numpy()
    got = rgb.cpu().numpy()
    for i in (0, 1, 12, 14, 19, 20, 21, 39):
        stream = slots[i, : int(lens[i])].cpu().numpy().tobytes()
        assert np.array_equal(got[i], ref.ref_decode(stream)), i
    assert not got[13].any()
    # packed form
    offs = np.concatenate([[0], np.cumsum(lens.astype(np.int64))])
    packed = torch.zeros(int(offs[-1]) + 64, dtype=torch.uint8, device="cuda")
    for i in range(40):
        packed[int(offs[i]): int(offs[i + 1])] = slots[i, : int(lens[i])]
    rgb2 = torch.empty_like(rgb)
    dst2 = torch.zeros_like(dst)
    codec.decode_packed_device(packed, torch.from_numpy(offs).cuda(), rgb2, dst2)
    assert torch.equal(dst2, dst) and torch.equal(rgb2, rgb)


def test_decode_planes_api(codec, ref):
    """nhw_decode_batch_planes: the Y/U/V byte planes decode_image leaves for write_image_bmp"""
    import ctypes
    imgs = _mixed(5, 9800)
    for q in (21, 12):
        streams = [ref.ref_encode(imgs[i], q) for i in range(5)]
        offs = np.zeros(6, dtype=np.uint64)
        offs[1:] = np.cumsum([len(s) for s in streams])
        blob = np.frombuffer(b"".join(streams), dtype=np.uint8)
        yuv = np.zeros((5, 786432), dtype=np.uint8)
        qual = np.zeros(5, dtype=np.int32)
        status = np.zeros(5, dtype=np.int32)
        rc = codec.lib.nhw_decode_batch_planes(codec.h, blob.ctypes.data, offs.ctypes.data, 5, yuv.ctypes.data,
                                               qual.ctypes.data, status.ctypes.data)
        assert rc == 0 and (status == 0).all() and (qual == q).all()
        for i in range(5):
            _, planes = ref.ref_decode(streams[i], planes=True)
            assert np.array_equal(yuv[i].reshape(3, 512, 512), planes), (q, i)


def test_decoder_color_fast_path_exhaustive(codec):
    """the back-end kernel's integer q >= 20 colour matrix == the IEEE double form of write_image_bmp on every one of
    the 2^24 (Y, U, V) triples"""
    assert codec.dec_color_check() == 0

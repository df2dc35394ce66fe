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

"""Batch I/O front-end (include/nhw_batchio.h), host side only: image readers for the BMP variants and PNM, 512x512 tiling,
the .nhwpack index parser against hostile files, and the exported symbols.  No GPU, no codec call."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

from nhwcodec_b200 import batchio as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_bmp(top_rgb, bpp=24, topdown=False, palette=None, bitfields=False):
    """top_rgb: (h, w, 3) RGB top-down (or (h, w) palette indices for bpp 8) -> BMP file bytes"""
    h, w = top_rgb.shape[:2]
    if bpp == 8:
        rows = top_rgb
    else:
        rows = top_rgb[:, :, ::-1]
        if bpp == 32:
            rows = np.concatenate([rows, np.full((h, w, 1), 200, np.uint8)], axis=2)
    if not topdown:
        rows = rows[::-1]
    rowbytes = (w * bpp // 8 + 3) & ~3
    data = b"".join(r.tobytes() + b"\0" * (rowbytes - w * bpp // 8) for r in rows)
    extra = b""
    if bpp == 8:
        extra = b"".join(bytes([c[2], c[1], c[0], 0]) for c in palette)
    if bitfields:
        extra = struct.pack("<III", 0x00FF0000, 0x0000FF00, 0x000000FF)
    off = 54 + len(extra)
    hd = b"BM" + struct.pack("<IHHI", off + len(data), 0, 0, off)
    hd += struct.pack("<IiiHHIIiiII", 40, w, -h if topdown else h, 1, bpp, 3 if bitfields else 0, len(data), 0, 0,
                      len(palette) if palette is not None else 0, 0)
    return hd + extra + data


@pytest.fixture(scope="module")
def picture():
    rng = np.random.default_rng(7)
    return rng.integers(0, 256, (300, 700, 3), dtype=np.uint8)    # RGB, top-down


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "nhw_batchio.h")).read()
    declared = set(re.findall(r"\b(nhw_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"nhw_tile_count"}          # mentioned in a comment only
    lib = ctypes.CDLL(B.LIB_PATH) if os.path.exists(B.LIB_PATH) else pytest.skip("libnhw_batchio.so not built")
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(B.EXPORTS), declared ^ set(B.EXPORTS)


@pytest.mark.parametrize("bpp", [24, 32])
@pytest.mark.parametrize("topdown", [False, True])
def test_bmp_variants_load_to_the_same_pixels(picture, bpp, topdown):
    want = picture[::-1, :, ::-1]            # bottom-up B,G,R: what read_image_bmp hands to the codec
    assert np.array_equal(B.load_image(make_bmp(picture, bpp, topdown)), want)


def test_bmp_bitfields_and_palette(picture):
    want = picture[::-1, :, ::-1]
    assert np.array_equal(B.load_image(make_bmp(picture, 32, False, bitfields=True)), want)
    rng = np.random.default_rng(3)
    pal = rng.integers(0, 256, (256, 3), dtype=np.uint8)
    idx = rng.integers(0, 256, (37, 41), dtype=np.uint8)           # odd width: row padding
    got = B.load_image(make_bmp(idx, 8, False, palette=[tuple(c) for c in pal]))
    assert np.array_equal(got, pal[idx][::-1, :, ::-1])


def test_pnm(picture):
    h, w = picture.shape[:2]
    ppm = b"P6\n# a comment\n%d %d\n255\n" % (w, h) + picture.tobytes()
    assert np.array_equal(B.load_image(ppm), picture[::-1, :, ::-1])
    grey = picture[:, :, 0]
    pgm = b"P5 %d %d 255\n" % (w, h) + grey.tobytes()
    assert np.array_equal(B.load_image(pgm), np.repeat(grey[::-1, :, None], 3, axis=2))


@pytest.mark.parametrize("bad", ["short", "rle", "bpp16", "truncated", "text", "ppm16"])
def test_rejects_unsupported_and_truncated(picture, bad):
    good = make_bmp(picture)
    data = {
        "short": good[:40],
        "rle": good[:30] + struct.pack("<I", 1) + good[34:],
        "bpp16": good[:28] + struct.pack("<H", 16) + good[30:],
        "truncated": good[:-1000],
        "text": b"hello world, not an image",
        "ppm16": b"P6 2 2 65535\n" + b"\0" * 24,
    }[bad]
    with pytest.raises(B.BatchIOError):
        B.load_image(data)


def test_tiling_round_trip_and_edge_replication(picture):
    img = picture[::-1, :, ::-1].copy()      # (300, 700): 2 x 1 tiles
    tiles = B.to_tiles(img)
    assert tiles.shape == (2, 786432)
    assert np.array_equal(B.from_tiles(tiles, 700, 300), img)
    t0 = tiles[0].reshape(512, 512, 3)       # bottom-up: tile row 511 is the picture's top line
    top = img[::-1]                          # picture rows from the top
    assert np.array_equal(t0[::-1][:300], top[:, :512])
    assert np.array_equal(t0[::-1][300:], np.repeat(top[299:300, :512], 212, axis=0))      # rows below the picture repeat its last line
    t1 = tiles[1].reshape(512, 512, 3)[::-1]
    assert np.array_equal(t1[:300, :188], top[:, 512:])
    assert np.array_equal(t1[:300, 188:], np.repeat(top[:, 699:700], 324, axis=1))         # columns right of it repeat its last column
    # a 512 x 512 picture is its own single tile, byte for byte
    sq = np.random.default_rng(1).integers(0, 256, (512, 512, 3), dtype=np.uint8)
    assert np.array_equal(B.to_tiles(sq)[0], sq.reshape(-1))


def test_save_load_round_trip(picture, tmp_path):
    img = picture[::-1, :, ::-1].copy()
    for fmt, ext in ((0, "bmp"), (1, "ppm")):
        p = str(tmp_path / ("a." + ext))
        B.save_image(p, img, fmt)
        assert np.array_equal(B.load_image(p), img)
    # for 512 x 512 the BMP header is the reference decoder's fixed one (decoder/nhw_decoder_cli.c:61-65)
    sq = np.zeros((512, 512, 3), dtype=np.uint8)
    p = str(tmp_path / "sq.bmp")
    B.save_image(p, sq, 0)
    want = bytes([66, 77, 54, 0, 12, 0, 0, 0, 0, 0, 54, 0, 0, 0, 40, 0, 0, 0, 0, 2, 0, 0, 0, 2, 0, 0, 1, 0, 24, 0, 0, 0, 0, 0, 0,
                  0, 12, 0] + [0] * 16)
    assert open(p, "rb").read(54) == want


def build_pack(blobs, images, quality=20):
    """the documented layout, written independently of the C writer.  images: (width, height, first_tile, name)"""
    body = b"NHWPACK1" + struct.pack("<II", 1, quality) + b"\0" * 16
    offs = []
    for b in blobs:
        offs.append(len(body))
        body += b
    offs.append(len(body))
    index_off = len(body)
    names = b""
    recs = b""
    for w, h, first, name in images:
        nb = name.encode()
        recs += struct.pack("<IIIIQII", w, h, (w + 511) // 512, (h + 511) // 512, first, len(names), len(nb))
        names += nb
    body += recs + b"".join(struct.pack("<Q", o) for o in offs) + names
    return body + struct.pack("<QQQ", index_off, len(images), len(blobs)) + b"NHWPKEND"


def test_pack_reader_against_the_documented_layout(tmp_path):
    blobs = [bytes([i]) * (10 + i) for i in range(3)]
    data = build_pack(blobs, [(512, 512, 0, "a.bmp"), (700, 300, 1, "dir/b.ppm")], quality=17)
    p = tmp_path / "x.nhwpack"
    p.write_bytes(data)
    with B.Pack(str(p)) as pk:
        assert (pk.n_images, pk.n_tiles, pk.quality) == (2, 3, 17)
        assert pk.image(1) == {"width": 700, "height": 300, "tiles_x": 2, "tiles_y": 1, "first_tile": 1, "name": "dir/b.ppm"}
        assert [pk.tile(t) for t in range(3)] == blobs
    out = tmp_path / "ex"
    st = B.extract_pack(str(p), str(out))
    assert st["tiles"] == 3
    assert (out / "a.nhw").read_bytes() == blobs[0]
    assert (out / "b.t0_0.nhw").read_bytes() == blobs[1] and (out / "b.t0_1.nhw").read_bytes() == blobs[2]


def test_pack_reader_rejects_hostile_files(tmp_path):
    blobs = [b"x" * 20, b"y" * 30]
    good = build_pack(blobs, [(512, 512, 0, "a"), (512, 512, 1, "b")])
    cases = {
        "truncated": good[:-5],
        "no_magic": b"NHWPACK2" + good[8:],
        "bad_trailer": good[:-8] + b"NHWPKENX",
        "index_past_end": good[:-32] + struct.pack("<QQQ", len(good), 2, 2) + b"NHWPKEND",
        "huge_counts": good[:-32] + struct.pack("<QQQ", 32, 1 << 50, 2) + b"NHWPKEND",
        "tile_outside": build_pack(blobs, [(512, 512, 0, "a"), (512, 512, 2, "b")]),            # first_tile + tiles > n_tiles
        "geometry_lies": build_pack(blobs, [(512, 512, 0, "a"), (2000, 512, 1, "b")]),          # 4 tiles claimed, 1 left
    }
    # offsets that run backwards
    idx = good.index(struct.pack("<Q", 32) + struct.pack("<Q", 52))
    cases["offsets_backwards"] = good[:idx] + struct.pack("<Q", 52) + struct.pack("<Q", 32) + good[idx + 16:]
    for name, data in cases.items():
        p = tmp_path / (name + ".nhwpack")
        p.write_bytes(data)
        with pytest.raises(B.BatchIOError):
            B.Pack(str(p))

"""Batch I/O front-end on the GPU: files -> .nhwpack -> files.
Parity bar as for the codec itself: every tile blob in the pack is byte for byte what the reference encoder writes for that
tile's pixels, every decoded image is what the reference decoder produces for those blobs (reassembled and cropped), and a
512 x 512 BMP comes back as exactly the file the reference's nhw-dec writes."""
import os
import subprocess

import numpy as np
import pytest

from nhwcodec_b200 import batchio as B
from nhwcodec_b200 import synth
from test_batchio_cpu import make_bmp
from test_oracle_cpu import bmp_header

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the reference decoder's fixed output header (decoder/nhw_decoder_cli.c:61-65)
DEC_HEADER = bytes([66, 77, 54, 0, 12, 0, 0, 0, 0, 0, 54, 0, 0, 0, 40, 0, 0, 0, 0, 2, 0, 0, 0, 2, 0, 0, 1, 0, 24, 0, 0, 0, 0, 0, 0, 0, 12, 0] + [0] * 16)


def top_rgb_of(pix):
    """786432 raw BMP pixel bytes (bottom-up B,G,R) -> (512, 512, 3) RGB top-down"""
    return np.ascontiguousarray(pix.reshape(512, 512, 3)[::-1, :, ::-1])


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    """a directory of mixed inputs; returns [(path, expected bottom-up BGR pixels)] in name order"""
    d = tmp_path_factory.mktemp("in")
    out = []
    a, b, c = synth.natural(5000), synth.natural(5001), synth.textured(5002)
    (d / "a_plain.bmp").write_bytes(bmp_header() + a.tobytes())
    out.append((str(d / "a_plain.bmp"), a.reshape(512, 512, 3)))
    (d / "b_topdown.bmp").write_bytes(make_bmp(top_rgb_of(b), 24, True))
    out.append((str(d / "b_topdown.bmp"), b.reshape(512, 512, 3)))
    (d / "c_32bit.bmp").write_bytes(make_bmp(top_rgb_of(c), 32, False))
    out.append((str(d / "c_32bit.bmp"), c.reshape(512, 512, 3)))
    # a 1030 x 600 picture cut from a 2 x 3 mosaic of generator images: 3 x 2 tiles, ragged on both edges
    mosaic = np.concatenate([np.concatenate([top_rgb_of(synth.natural(5010 + 3 * r + k)) for k in range(3)], axis=1) for r in range(2)], axis=0)
    big = np.ascontiguousarray(mosaic[:600, :1030])
    h, w = big.shape[:2]
    (d / "d_big.ppm").write_bytes(b"P6\n%d %d\n255\n" % (w, h) + big.tobytes())
    out.append((str(d / "d_big.ppm"), np.ascontiguousarray(big[::-1, :, ::-1])))
    small = np.ascontiguousarray(top_rgb_of(a)[:200, :333])
    (d / "e_small.bmp").write_bytes(make_bmp(small, 24, False))       # 333 * 3 = 999 bytes per row: padded rows
    out.append((str(d / "e_small.bmp"), np.ascontiguousarray(small[::-1, :, ::-1])))
    (d / "notes.txt").write_text("not an image; the directory scan skips it\n")
    return str(d), out


def check_pack(pack_path, files, ref, q):
    """every blob == the reference encoder on that tile; returns the expected decoded pixels per image"""
    expected = []
    with B.Pack(pack_path) as pk:
        assert pk.n_images == len(files) and pk.quality == q
        for i, (path, pixels) in enumerate(files):
            info = pk.image(i)
            h, w = pixels.shape[:2]
            assert (info["width"], info["height"], info["name"]) == (w, h, path)
            tiles = B.to_tiles(pixels)
            assert info["tiles_x"] * info["tiles_y"] == tiles.shape[0]
            dec = np.empty_like(tiles)
            for t in range(tiles.shape[0]):
                blob = pk.tile(info["first_tile"] + t)
                assert blob == ref.ref_encode(tiles[t], q), (path, t)
                dec[t] = ref.ref_decode(blob)
            expected.append(B.from_tiles(dec, w, h))
    return expected


@pytest.mark.parametrize("q,group", [(20, 0), (9, 3)])
def test_directory_to_pack_to_directory(files, codec, ref, tmp_path, q, group):
    """group 3: the five images (1 + 1 + 1 + 6 + 1 tiles) do not fit one staging buffer -> carry-over between the two slots,
    the 6-tile image forces... a buffer of its own size"""
    d, items = files
    pack = str(tmp_path / "set.nhwpack")
    if group and group < 6:
        with pytest.raises(B.BatchIOError):          # the 6-tile picture cannot fit 3-tile staging buffers: reported, not truncated
            B.encode_dir(codec, d, q, pack, group)
        group = 6
    st = B.encode_dir(codec, d, q, pack, group)
    assert (st["images"], st["tiles"], st["first_bad_image"]) == (5, 10, -1)
    assert st["bytes_out"] == os.path.getsize(pack)
    expected = check_pack(pack, items, ref, q)
    out = tmp_path / "out"
    st = B.decode_pack(codec, pack, str(out), 0, group)
    assert (st["images"], st["tiles"]) == (5, 10)
    for (path, _), want in zip(items, expected):
        stem = os.path.splitext(os.path.basename(path))[0]
        assert np.array_equal(B.load_image(str(out / (stem + ".bmp"))), want), path
    # the 512 x 512 ones are exactly the files the reference decoder writes: fixed header + pixels
    got = (out / "a_plain.bmp").read_bytes()
    assert got == DEC_HEADER + expected[0].tobytes()
    # PPM output carries the same pixels
    out2 = tmp_path / "out_ppm"
    B.decode_pack(codec, pack, str(out2), 1, group)
    assert np.array_equal(B.load_image(str(out2 / "d_big.ppm")), expected[3])


def test_manifest_cli_and_extract_interop(files, ref, tmp_path):
    """nhw-batch (the CLI) driven by a manifest; the extracted tiles are ordinary .nhw files: the reference's own decoder
    CLI source (linked against our library) and ours read them"""
    d, items = files
    exe = os.path.join(ROOT, "cli", "nhw-batch")
    if not os.path.exists(exe):
        pytest.skip("cli/nhw-batch not built")
    man = tmp_path / "list.txt"
    man.write_text("# two of the five\n\n%s\n  %s  \n" % (items[1][0], items[3][0]))
    pack = str(tmp_path / "m.nhwpack")
    r = subprocess.run([exe, "enc", "-q20", "-g8", "-o", pack, "-m", str(man)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    check_pack(pack, [items[1], items[3]], ref, 20)
    r = subprocess.run([exe, "list", pack], capture_output=True, text=True)
    assert r.returncode == 0 and "2 images, 7 tiles, quality 20" in r.stdout
    ex = tmp_path / "ex"
    r = subprocess.run([exe, "extract", pack, str(ex)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    nhw = ex / "b_topdown.nhw"
    assert nhw.read_bytes() == ref.ref_encode(items[1][1].reshape(-1), 20)
    bmp = str(tmp_path / "b.bmp")
    r = subprocess.run([os.path.join(ROOT, "cli", "nhw-dec"), str(nhw), bmp], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(bmp, "rb").read() == DEC_HEADER + ref.ref_decode(nhw.read_bytes()).tobytes()
    assert len(list(ex.glob("d_big.t*_*.nhw"))) == 6
    r = subprocess.run([exe, "dec", pack, str(tmp_path / "dec")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "dec" / "d_big.bmp").exists()


def test_missing_file_stops_the_job_and_keeps_what_was_done(files, codec, tmp_path):
    d, items = files
    pack = str(tmp_path / "p.nhwpack")
    paths = [items[0][0], items[2][0], str(tmp_path / "does_not_exist.bmp"), items[1][0]]
    with pytest.raises(B.BatchIOError) as e:
        B.encode_files(codec, paths, 20, pack)
    assert "does_not_exist" in str(e.value)
    with B.Pack(pack) as pk:               # a valid pack of the images before the failure
        assert pk.n_images == 2 and pk.n_tiles == 2

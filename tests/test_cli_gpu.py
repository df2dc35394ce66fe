"""Drop-in boundary on the GPU: our nhw-enc and the reference's own CLI source linked against
libnhw_compat/libnhw_cuda must write the canonical .nhw bytes."""
import hashlib
import os
import subprocess

import pytest

from test_oracle_cpu import bmp_header, smooth_pixels

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT_Q20 = "9ea5053bd20fc35758652b0a0aa47b8d"


@pytest.fixture(scope="module")
def smooth_bmp(tmp_path_factory):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    p = tmp_path_factory.mktemp("cli") / "smooth.bmp"
    p.write_bytes(bmp_header() + smooth_pixels().tobytes())
    return str(p)


@pytest.mark.parametrize("exe", ["cli/nhw-enc", "oracle/_ref/nhw-enc-dropin"])
def test_cli_known_answer(smooth_bmp, exe, tmp_path):
    path = os.path.join(ROOT, exe)
    if not os.path.exists(path):
        pytest.skip(exe + " not built")
    out = str(tmp_path / "out.nhw")
    r = subprocess.run([path, "-q20", smooth_bmp, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    data = open(out, "rb").read()
    assert len(data) == 24412
    assert hashlib.md5(data).hexdigest() == KAT_Q20
    # refuses to overwrite without -f, like the reference (encoder/nhw_encoder_cli.c:163-172)
    r2 = subprocess.run([os.path.join(ROOT, "cli/nhw-enc"), "-q20", smooth_bmp, out], capture_output=True, text=True)
    assert r2.returncode == 1
    r3 = subprocess.run([os.path.join(ROOT, "cli/nhw-enc"), "-f", "-q20", smooth_bmp, out], capture_output=True, text=True)
    assert r3.returncode == 0


@pytest.mark.parametrize("exe", ["cli/nhw-enc", "oracle/_ref/nhw-enc-dropin"])
@pytest.mark.parametrize("q", [17, 23])
def test_cli_other_qualities(smooth_bmp, exe, q, tmp_path):
    """-q17 (scaled colour path, no res3/res5) and -q23 (res6 / char_res1 / high_qsetting3 sections) through both
    CLIs, then back through both decoder CLIs: SURVEY.md Appendix E hashes"""
    from test_oracle_cpu import KAT
    path = os.path.join(ROOT, exe)
    if not os.path.exists(path):
        pytest.skip(exe + " not built")
    out = str(tmp_path / "out.nhw")
    r = subprocess.run([path, "-q%d" % q, smooth_bmp, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    data = open(out, "rb").read()
    assert (len(data), hashlib.md5(data).hexdigest()) == (KAT[q][0], KAT[q][1])
    for dec in ("cli/nhw-dec", "oracle/_ref/nhw-dec-dropin"):
        dpath = os.path.join(ROOT, dec)
        if not os.path.exists(dpath):
            continue
        bmp = str(tmp_path / "back.bmp")
        r = subprocess.run([dpath, out, bmp], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert hashlib.md5(open(bmp, "rb").read()).hexdigest() == KAT[q][2], dec


def test_cli_rejects_bad_input(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    bad = tmp_path / "bad.bmp"
    bad.write_bytes(b"XX" + b"\0" * 100)
    r = subprocess.run([os.path.join(ROOT, "cli/nhw-enc"), str(bad), str(tmp_path / "o.nhw")], capture_output=True)
    assert r.returncode == (-13) % 256     # HEADER_CHECK_ERROR_NO_VALID_SIG, encoder/nhw_encoder.c:63-71
    r = subprocess.run([os.path.join(ROOT, "cli/nhw-enc"), "-q99", "a", "b"], capture_output=True)
    assert r.returncode == 1


KAT_DEC_Q20 = "1d9ccf4e698182fcac1ae31992789caf"   # nhw-dec output of the canonical q20 stream


@pytest.mark.parametrize("exe", ["cli/nhw-dec", "oracle/_ref/nhw-dec-dropin"])
def test_decoder_cli_known_answer(smooth_bmp, exe, tmp_path):
    """our nhw-dec, and the reference's own decoder CLI source linked against libnhw_compat_dec,
    must write the BMP the reference nhw-dec writes"""
    path = os.path.join(ROOT, exe)
    if not os.path.exists(path):
        pytest.skip(exe + " not built")
    nhw = str(tmp_path / "s.nhw")
    r = subprocess.run([os.path.join(ROOT, "cli/nhw-enc"), "-q20", smooth_bmp, nhw], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = str(tmp_path / "out.bmp")
    r = subprocess.run([path, nhw, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    data = open(out, "rb").read()
    assert len(data) == 786486
    assert hashlib.md5(data).hexdigest() == KAT_DEC_Q20


def test_cli_top_down_bmp_and_q0(smooth_bmp, tmp_path):
    """a negative-height (top-down) BMP is flipped before encoding, like the reference does (encoder/nhw_encoder.c:3089-3093):
    the same picture stored both ways gives the same .nhw; -q0 -- which the reference accepts although its tables are then
    indexed out of bounds -- is refused with the exit code of an encoder failure"""
    import struct
    data = open(smooth_bmp, "rb").read()
    hd, pix = bytearray(data[:54]), data[54:]
    hd[22:26] = struct.pack("<i", -512)
    rows = [pix[i * 1536:(i + 1) * 1536] for i in range(512)]
    td = tmp_path / "topdown.bmp"
    td.write_bytes(bytes(hd) + b"".join(reversed(rows)))
    out = str(tmp_path / "td.nhw")
    for exe in ("cli/nhw-enc", "oracle/_ref/nhw-enc-dropin"):
        path = os.path.join(ROOT, exe)
        if not os.path.exists(path):
            continue
        r = subprocess.run([path, "-f", "-q20", str(td), out], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert hashlib.md5(open(out, "rb").read()).hexdigest() == KAT_Q20, exe
    canon = os.path.join(ROOT, "oracle/_ref/nhw-enc-canon")      # the reference itself on the same file
    if os.path.exists(canon):
        ref_out = str(tmp_path / "td_ref.nhw")
        assert subprocess.run([canon, "-f", "-q20", str(td), ref_out], capture_output=True).returncode == 0
        assert open(ref_out, "rb").read() == open(out, "rb").read()
    r = subprocess.run([os.path.join(ROOT, "cli/nhw-enc"), "-f", "-q0", smooth_bmp, out], capture_output=True, text=True)
    assert r.returncode == (-1) % 256 and "encode failed" in r.stderr

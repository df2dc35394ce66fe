"""The encoder / decoder stage functions (the same __host__ __device__ sources the kernels are built from),
compiled for the host and run in the kernels' schedules (wavefront steps, segments, pointwise forms) by
tests/hostemu, against the oracle.  Test tooling around product code: catches a wrong dependency analysis
without a GPU.  The GPU parity tests (tests/test_*_gpu.py) remain the parity claim."""
import os
import subprocess
import sys

import pytest

from nhwcodec_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HE = os.path.join(ROOT, "tests", "hostemu")


@pytest.fixture(scope="module")
def he(ref):
    subprocess.check_call(["bash", os.path.join(HE, "build.sh")])
    sys.path.insert(0, HE)
    import run as he_run
    return he_run


@pytest.mark.parametrize("kind,seed,q", [("natural", 1000, 20), ("noise", 5, 23), ("textured", 1002, 17), ("natural", 7, 18),
                                          ("textured", 11, 21), ("natural", 3, 22), ("noise", 9, 19)])
def test_host_schedule_encode(he, kind, seed, q):
    pix = getattr(synth, kind)(seed)
    ok, first_bad, stream, ref_stream = he.compare(pix, q, verbose=False)
    assert ok, (first_bad, len(stream) if isinstance(stream, bytes) else stream, len(ref_stream))


@pytest.mark.parametrize("kind,seed,q", [("natural", 1000, 20), ("noise", 5, 22), ("textured", 1002, 18), ("noise", 7, 12),
                                          ("textured", 21, 15), ("natural", 4, 3)])
def test_host_schedule_decode(he, kind, seed, q):
    assert he.compare_decode(getattr(synth, kind)(seed), q, verbose=False)


@pytest.mark.parametrize("kind,seed,q", [("natural", 1000, 16), ("textured", 1002, 14), ("noise", 7, 13), ("natural", 12, 12),
                                          ("textured", 13, 10), ("noise", 14, 7), ("natural", 15, 5), ("textured", 16, 1)])
def test_host_schedule_encode_low_quality(he, kind, seed, q):
    """q <= 16: the row / image forms the CUDA path runs at these settings (E7, E8, low E14, cyclic quantisers, chroma
    pre-filter, thresholds, LL smoothing), from the reference's own front-end planes"""
    pix = getattr(synth, kind)(seed)
    ok, first_bad, stream, ref_stream = he.compare(pix, q, verbose=False)
    assert ok, (first_bad, len(stream) if isinstance(stream, bytes) else stream, len(ref_stream))


@pytest.mark.parametrize("q", [16, 15, 14, 11, 9, 8, 6, 2])
def test_pre_sharpening_state_machine(he, ref, q):
    """pre_lowq.cuh against the reference's pre_processing stage on the reference's own luma plane"""
    import ctypes
    import numpy as np
    L = he.lib()
    L.he_pre_lowq.argtypes = [ctypes.c_void_p, ctypes.c_int]
    for kind, seed in (("natural", 1000 + q), ("textured", 2000 + q), ("noise", 3000 + q)):
        Y, _, _ = ref.ref_colorspace(getattr(synth, kind)(seed), q)
        want = ref.ref_pre_processing(Y, q)
        got = np.ascontiguousarray(Y).copy()
        L.he_pre_lowq(got.ctypes.data, q)
        assert np.array_equal(got, want), (q, kind, int((got != want).sum()))


def test_hostile_headers_are_rejected_on_the_host(he, ref):
    """nhw_parse_header (dec_parse.h) is the gate in front of the decode kernels: streams whose section lengths
    exceed the decode workspace, truncated streams and lying length fields must be refused before any kernel
    sees them"""
    import struct
    good = ref.ref_encode(synth.natural(31), 19)
    rc, _, _ = he.host_decode(good)
    assert rc == 0
    cases = []
    b = bytearray(good); struct.pack_into("<H", b, 22, 65535); cases.append(bytes(b) + b"\0" * 300000)   # res3_bit_len
    b = bytearray(good); struct.pack_into("<H", b, 14, 65535); cases.append(bytes(b))                     # tree_end
    b = bytearray(good); struct.pack_into("<H", b, 26, 60000); cases.append(bytes(b) + b"\0" * 300000)   # res1_bit_len
    cases += [good[:36], good[:20], good[:1], b""]
    b = bytearray(good); b[1] = 23; cases.append(bytes(b[:37]))
    b = bytearray(good); b[1] = 0; cases.append(bytes(b))
    b = bytearray(good); b[0] = 9; cases.append(bytes(b))
    b = bytearray(good); struct.pack_into("<I", b, 10, 1 << 30); cases.append(bytes(b))                   # size_data2
    for k, s in enumerate(cases):
        rc, _, _ = he.host_decode(s)
        assert rc != 0, k
    # a res4 list that claims 65535 entries (the sections behind it then land in the zero padding): nothing the
    # header check can object to, so the walk itself has to stay inside the LL2 band
    b = bytearray(good); struct.pack_into("<H", b, 24, 65535)
    he.host_decode(bytes(b) + b"\0" * 300000)


def test_decoder_fuzz_under_asan(he, ref):
    """mutated streams through the decoder stage functions built with AddressSanitizer (tests/hostemu/fuzz_decode.py)"""
    out = subprocess.run([sys.executable, os.path.join(HE, "fuzz_decode.py"), "11", "12"], capture_output=True, text=True)
    assert out.returncode == 0 and "no memory error" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]

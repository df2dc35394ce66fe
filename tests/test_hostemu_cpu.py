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


@pytest.mark.parametrize("kind,seed,q", [("natural", 1000, 20), ("noise", 5, 22), ("textured", 1002, 18)])
def test_host_schedule_decode(he, kind, seed, q):
    assert he.compare_decode(getattr(synth, kind)(seed), q, verbose=False)

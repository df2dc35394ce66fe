import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref():
    """The canonicalised compiled reference (oracle/_ref).  Test infrastructure only."""
    from oracle import refbind
    if not refbind.available():
        import subprocess
        subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")])
    if not refbind.available():
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    return refbind


@pytest.fixture(scope="session")
def codec():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from nhwcodec_b200 import Codec
    c = Codec(device=0, max_batch=16)
    yield c
    c.close()

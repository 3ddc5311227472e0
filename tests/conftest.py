import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The product library and the C oracle must exist before any test (built by `make` /
    __graft_entry__.build(); here as a convenience for a bare checkout)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "gpu_sdr_b200", "libgsdr.so")):
        subprocess.check_call(["make", "-C", root, "-j8"])
    from oracle import gsdr_oracle
    gsdr_oracle.build()


@pytest.fixture(scope="session")
def gpu_required():
    from common import has_gpu
    if not has_gpu():
        pytest.skip("no CUDA device visible")

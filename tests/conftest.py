import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session", autouse=True)
def _forward_path_override():
    """DCB_TEST_FWD_PATH=1|2 runs the whole suite with the forward pinned to one kernel family
    (1 = round-1 accumulator pipeline, 2 = target-tile owner); default: the library's own dispatch."""
    v = os.environ.get("DCB_TEST_FWD_PATH")
    if v:
        import diffcodec_b200
        diffcodec_b200._lib.set_option("fwd_path", int(v))
    yield

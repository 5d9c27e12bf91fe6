"""pytest config: `gpu` marker (tests that need a real B200) and repo-root imports."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _drain_device_between_tests():
    """GVL_TEST_SYNC=1 (tool runs): drain the device and collect garbage after every test, so that an asynchronous error is
    reported by the test that caused it and no object dies with device work in flight."""
    yield
    import os

    if os.environ.get("GVL_TEST_SYNC") == "1":
        import gc

        import torch

        if torch.cuda.is_available():
            torch.cuda.synchronize()
            gc.collect()
            torch.cuda.synchronize()

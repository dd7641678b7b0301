import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _two_envs_per_warp_for_small_test_batches(monkeypatch):
    """Batches below 1,536 envs normally take the latency-optimised one-env-per-warp mapping of the Transport / RD
    kernels.  The parity tests run small batches, so by default they force the throughput mapping (two envs per warp on
    the reference grids: the headline kernel); tests of the small-batch mapping override the variable themselves."""
    monkeypatch.setenv("PDECG_BC1D_HALFWARP_MIN", "0")

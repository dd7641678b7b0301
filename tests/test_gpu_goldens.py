"""GPU: the CUDA envs replay the committed runs of the unmodified reference (tests/golden/*.npz)."""
import pytest

from tests import golden_replay as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", R.bc1d_cases())
def test_bc1d_reference_runs(cuda_device, name):
    R.replay_bc1d(name, "cuda", cuda_device, reward_rtol=2e-6)


@pytest.mark.parametrize("key", R.arz_cases())
def test_arz_reference_runs(cuda_device, key):
    R.replay_arz(key, "cuda", cuda_device)


def test_brain_tumor_raw_reference_run(cuda_device):
    R.replay_brain_tumor_raw("cuda", cuda_device)


@pytest.mark.parametrize("name", ["open", "paper", "paper_weekends"])
def test_brain_tumor_wrapper_reference_runs(cuda_device, name):
    R.replay_brain_tumor_wrapper(name, "cuda", cuda_device)


def test_ns_reference_run(cuda_device):
    R.replay_ns("cuda", cuda_device)

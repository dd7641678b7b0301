"""GPU: the device exp_cr() equals its plain-C restatement bit for bit (which tests/test_exp_cr.py pins to mpmath)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_exp_cr_matches_the_c_restatement(cuda_device):
    import torch
    from pdecontrolgym_b200 import _lib
    from tests.test_exp_cr import exp_cr
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(-17.5, 0, 200000), rng.uniform(-745, 709.7, 100000), rng.uniform(-1e-6, 1e-6, 1000),
                        np.array([0.0, -0.0, 709.79, -745.2, np.inf, -np.inf, 1.0, -1.0])])
    xd = torch.from_numpy(x).to(cuda_device)
    out = torch.empty_like(xd)
    _lib.check(_lib.lib().pdecg_selftest_exp(x.size, xd.data_ptr(), out.data_ptr(), _lib.current_raw_stream(cuda_device.index)))
    got = out.cpu().numpy()
    want = exp_cr(x)
    sub = want < 2.3e-308                    # subnormal results: rounded twice on both sides, identically
    assert np.array_equal(got, want), f"{int((got != want).sum())} mismatches (subnormal: {int((got != want)[sub].sum())})"
    nan = torch.full((4,), float("nan"), dtype=torch.float64, device=cuda_device)
    _lib.check(_lib.lib().pdecg_selftest_exp(4, nan.data_ptr(), out.data_ptr(), _lib.current_raw_stream(cuda_device.index)))
    assert torch.isnan(out[:4]).all()

"""GPU: the kernels' division-by-a-loop-invariant-divisor helper (csrc/pdecg_common.cuh: ConstDiv) must be
bit-identical to IEEE division for EVERY input -- it is what keeps Transport / BrainTumor / ARZ states
bit-exact while removing the reciprocal computation from the inner loops."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs(dtype, n, rng):
    info = np.finfo(dtype)
    bits = np.uint32 if dtype == np.float32 else np.uint64
    nb = 32 if dtype == np.float32 else 64
    # uniformly random bit patterns (all exponents, denormals, inf/nan included) + ordinary magnitudes + edge values
    raw = rng.integers(0, 2 ** nb, size=n // 2, dtype=np.uint64).astype(bits).view(dtype)
    ordinary = (rng.standard_normal(n // 2) * 10.0 ** rng.uniform(-6, 6, n // 2)).astype(dtype)
    edge = np.array([0.0, -0.0, info.tiny, -info.tiny, info.tiny / 4, info.max, -info.max, np.inf, -np.inf, np.nan,
                     1.0, -1.0, info.eps, 1 / 3, 2.0 ** -80, 2.0 ** 81, 2.0 ** -81, 2.0 ** 80], dtype=dtype)
    return np.concatenate([raw, ordinary, edge])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("d", [0.01, 0.005, 1.0, 1.0 / 3.0, 1e5, 0.16, 2.5e-5, 0.1, 7.0, 2.0 ** -40, 2.0 ** 40, 1e-300, 3e37])
def test_constdiv_bit_identical_to_ieee_division(cuda_device, dtype, d):
    import torch
    from pdecontrolgym_b200 import _lib
    if dtype == np.float32 and not (1e-38 < abs(d) < 3e38):
        pytest.skip("divisor not representable in float32")
    rng = np.random.default_rng(int(abs(d) * 1000) % 9973 + (0 if dtype == np.float32 else 1))
    x = _inputs(dtype, 1 << 21, rng)
    xt = torch.from_numpy(x).to(cuda_device)
    fast, ref = torch.empty_like(xt), torch.empty_like(xt)
    L = _lib.lib()
    _lib.check(L.pdecg_selftest_constdiv(_lib.F32 if dtype == np.float32 else _lib.F64, xt.numel(), xt.data_ptr(), C.c_double(d),
                                         fast.data_ptr(), ref.data_ptr(), torch.cuda.current_stream(cuda_device).cuda_stream))
    f, r = fast.cpu().numpy(), ref.cpu().numpy()
    with np.errstate(all="ignore"):
        want = (x / dtype(d)).astype(dtype)
    ib = np.uint32 if dtype == np.float32 else np.uint64
    nan = np.isnan(r)
    assert np.array_equal(np.isnan(f), nan)
    assert np.array_equal(f[~nan].view(ib), r[~nan].view(ib)), "ConstDiv differs from __fdiv_rn/__ddiv_rn"
    assert np.array_equal(r[~nan].view(ib), want[~nan].view(ib)), "device IEEE division differs from numpy"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fast_div_bit_identical_inside_its_window(cuda_device, dtype):
    """The deferred-check division of the Traffic-ARZ / BrainTumor inner loops (csrc/pdecg_common.cuh: fast_div +
    DivWindow): wherever the window predicate holds, the straight-line quotient equals IEEE division bit for bit;
    the window must cover the magnitudes the PDE states take (so the fallback is the rare path) and must reject
    zeros, denormals, infinities and NaNs."""
    import torch
    from pdecontrolgym_b200 import _lib
    rng = np.random.default_rng(7 if dtype == np.float32 else 8)
    n = 1 << 22
    a = _inputs(dtype, n, rng)
    b = _inputs(dtype, n, np.random.default_rng(11))
    rng.shuffle(b)
    # a third of the pairs: ordinary PDE magnitudes (densities 0.05..0.2, momenta -1..1)
    m = a.size // 3
    a[:m] = rng.uniform(-1, 1, m).astype(dtype)
    b[:m] = rng.uniform(0.05, 0.2, m).astype(dtype)
    at, bt = torch.from_numpy(a).to(cuda_device), torch.from_numpy(b).to(cuda_device)
    fast, ref = torch.empty_like(at), torch.empty_like(at)
    ok = torch.empty(at.numel(), dtype=torch.uint8, device=cuda_device)
    L = _lib.lib()
    _lib.check(L.pdecg_selftest_fastdiv(_lib.F32 if dtype == np.float32 else _lib.F64, at.numel(), at.data_ptr(), bt.data_ptr(),
                                        fast.data_ptr(), ok.data_ptr(), ref.data_ptr(),
                                        torch.cuda.current_stream(cuda_device).cuda_stream))
    f, r, k = fast.cpu().numpy(), ref.cpu().numpy(), ok.cpu().numpy().astype(bool)
    ib = np.uint32 if dtype == np.float32 else np.uint64
    with np.errstate(all="ignore"):
        want = (a / b).astype(dtype)
    assert k[:m].all(), "ordinary magnitudes must stay on the fast path"
    assert k.mean() > 0.3
    assert not np.isnan(r[k]).any() and np.isfinite(r[k]).all()
    assert np.array_equal(f[k].view(ib), r[k].view(ib)), "fast_div differs from the device's IEEE division inside its window"
    assert np.array_equal(r[k].view(ib), want[k].view(ib))
    bad = (a == 0) | (b == 0) | ~np.isfinite(a) | ~np.isfinite(b) | (np.abs(a) < np.finfo(dtype).tiny) | (np.abs(b) < np.finfo(dtype).tiny)
    assert not k[bad].any(), "zeros / denormals / inf / NaN must be routed to the exact division"

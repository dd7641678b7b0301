"""GPU: Philox-keyed on-device initial conditions (pdecg_bc1d_reset_random): values equal a numpy
restatement of Philox4x32-10, are invariant to sharding, and change per episode under autoreset."""
import numpy as np
import pytest

from tests.helpers import bc1d_case

pytestmark = pytest.mark.gpu
M32 = 0xFFFFFFFF


def philox_uniform53(seed, a, b):
    c = [a & M32, (a >> 32) & M32, b & M32, 0]
    k0, k1 = seed & M32, (seed >> 32) & M32
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k1) & M32, p0 & M32]
        k0, k1 = (k0 + 0x9E3779B9) & M32, (k1 + 0xBB67AE85) & M32
    return float(((c[0] << 32) | c[1]) >> 11) * (1.0 / 9007199254740992.0)


def test_values_sharding_and_autoreset(cuda_device):
    import torch
    import pdecontrolgym_b200 as pcg
    kw, npts, beta, nt = bc1d_case("rd", None, 2)
    N, seed, lo, hi = 11, 1234567890123, 1.0, 10.0
    prof = np.linspace(0.5, 1.5, npts)
    mk = lambda n, off: pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=n, device=cuda_device, precision="fp64",  # noqa: E731
                                 reward_class=pcg.NormReward(nt - 1), normalize=True, autoreset=True, env_id_offset=off, **kw)
    env = mk(N, 0)
    obs, _ = env.reset(seed=seed, options={"init_condition": pcg.UniformScaleIC(lo, hi, prof), "beta": beta})
    want = np.stack([(lo + (hi - lo) * philox_uniform53(seed, e, 0)) * prof for e in range(N)])
    assert np.array_equal(obs.cpu().numpy(), want)
    # two shards reproduce the single-batch draws
    a, b = mk(4, 0), mk(7, 4)
    oa, _ = a.reset(seed=seed, options={"init_condition": pcg.UniformScaleIC(lo, hi, prof), "beta": beta})
    ob, _ = b.reset(seed=seed, options={"init_condition": pcg.UniformScaleIC(lo, hi, prof), "beta": beta})
    assert np.array_equal(np.concatenate([oa.cpu().numpy(), ob.cpu().numpy()]), want)
    # episode 2 after same-step autoreset: fresh draws with episode counter 1
    for k in range(2):
        obs, r, te, tr, _ = env.step(torch.zeros(N))
    assert te.all()
    want2 = np.stack([(lo + (hi - lo) * philox_uniform53(seed, e, 1)) * prof for e in range(N)])
    assert np.array_equal(obs.cpu().numpy(), want2)
    assert (env.time_index == 0).all()

"""Known-answer tests K1 / K2: the reference's PUBLISHED episode rewards and summed state norms for the
backstepping-controlled Transport and ReactionDiffusion envs (HyperbolicPDEExample.ipynb:820,823 and
ParabolicPDEExample.ipynb:33152,33155; also docs/source/_static/img/hyperbolicTable.png), u(x,0) = 1 and 10.

Tolerance 1e-6 relative: the notebooks ran numpy 1.26 where Python-side reward accumulation promoted to
float64; under numpy >= 2 (NEP 50) the reference itself reproduces these values only to float32 precision
(SURVEY.md 8(c) "version caveat").  The controller grid is np.linspace(dx, X, round(X/dx)) as in
transport1Dbackstepping.py:88-89 (the notebook's displayed `+1` grid does not reproduce its own stored output)."""
import numpy as np
import pytest

from oracle import np_oracle as O
from tests import kat_controllers as KC

RTOL = 1e-6


def _setup(env):
    if env == "tr":
        dx, npts, gamma, amp = 1e-2, 100, 7.35, 5
        kern = KC.transport_kernel(KC.chebyshev_beta(np.linspace(dx, 1, 100), gamma, amp), dx)
        control = lambda o: KC.transport_control(kern, o)  # noqa: E731
        return KC.TRANSPORT_KW, npts, KC.chebyshev_beta(np.linspace(0, 1, npts), gamma, amp), control, 50000, KC.K1, 50
    dx, npts, gamma, amp = 5e-3, 201, 8, 50
    kern = KC.rd_kernel(KC.chebyshev_beta(np.linspace(dx, 1, 200), gamma, amp), dx)
    control = lambda o: KC.rd_control(kern, o, dx)  # noqa: E731
    return KC.RD_KW, npts, KC.chebyshev_beta(np.linspace(0, 1, npts), gamma, amp), control, 100000, KC.K2, 1000


@pytest.mark.parametrize("env", ["tr", "rd"])
@pytest.mark.parametrize("u0", [1, 10])
def test_oracle_reproduces_published_values(env, u0):
    kw, npts, beta, control, nt_r, want, n_steps = _setup(env)
    Orc = O.TransportOracle if env == "tr" else O.ReactionDiffusionOracle
    e = Orc(1, reward=O.TunedReward1DSpec(nt_r, -1e3, 3e2), **kw)
    total, norm_sum, n = KC.run_episode(lambda: e.reset(np.ones(npts) * u0, beta)[0],
                                        lambda a: tuple(x[0] for x in e.step(np.array([a]))), control)
    assert n == n_steps
    assert abs(total - want[u0][0]) <= RTOL * abs(want[u0][0]), (total, want[u0][0])
    assert abs(norm_sum - want[u0][1]) <= RTOL * abs(want[u0][1]), (norm_sum, want[u0][1])


@pytest.mark.gpu
@pytest.mark.parametrize("env", ["tr", "rd"])
def test_cuda_reproduces_published_values(cuda_device, env):
    """Both initial conditions as a batch of 2 envs; the controller runs on the host per env."""
    import pdecontrolgym_b200 as pcg
    kw, npts, beta, control, nt_r, want, n_steps = _setup(env)
    eid = "PDEControlGym-TransportPDE1D" if env == "tr" else "PDEControlGym-ReactionDiffusionPDE1D"
    genv = pcg.make(eid, num_envs=2, device=cuda_device, precision="ref", reward_class=pcg.TunedReward1D(nt_r, -1e3, 3e2), **kw)
    ic = np.stack([np.ones(npts) * 1.0, np.ones(npts) * 10.0])
    obs, _ = genv.reset(options={"init_condition": ic, "beta": beta})
    obs = obs.cpu().numpy()
    total, norm_sum, n = np.zeros(2), np.zeros(2), 0
    done = False
    while not done:
        a = np.array([control(obs[0]), control(obs[1])], dtype=np.float64)
        o, r, te, tr, _ = genv.step(a)
        obs = o.cpu().numpy()
        total += r.cpu().numpy()
        norm_sum += np.array([np.linalg.norm(obs[0]), np.linalg.norm(obs[1])])
        n += 1
        done = bool(te.all()) or bool(tr.any())
    assert n == n_steps
    for i, u0 in enumerate((1, 10)):
        assert abs(total[i] - want[u0][0]) <= RTOL * abs(want[u0][0]), (u0, total[i], want[u0][0])
        assert abs(norm_sum[i] - want[u0][1]) <= RTOL * abs(want[u0][1]), (u0, norm_sum[i], want[u0][1])


@pytest.mark.gpu
@pytest.mark.parametrize("env", ["tr", "rd"])
def test_cuda_closed_loop_on_device(cuda_device, env):
    """Same published values with the batched device-side controllers (no host round trip per step) on a batch
    that tiles the two initial conditions 8 times."""
    import torch
    import pdecontrolgym_b200 as pcg
    from pdecontrolgym_b200.controllers import ReactionDiffusionBackstepping, TransportBackstepping
    kw, npts, beta, _, nt_r, want, n_steps = _setup(env)
    if env == "tr":
        ctl = TransportBackstepping(KC.chebyshev_beta(np.linspace(1e-2, 1, 100), 7.35, 5), 1e-2, cuda_device)
        eid = "PDEControlGym-TransportPDE1D"
    else:
        ctl = ReactionDiffusionBackstepping(KC.chebyshev_beta(np.linspace(5e-3, 1, 200), 8, 50), 5e-3, cuda_device)
        eid = "PDEControlGym-ReactionDiffusionPDE1D"
    N = 16
    genv = pcg.make(eid, num_envs=N, device=cuda_device, precision="ref", reward_class=pcg.TunedReward1D(nt_r, -1e3, 3e2), copy_obs=False, **kw)
    ic = np.tile(np.stack([np.ones(npts) * 1.0, np.ones(npts) * 10.0]), (N // 2, 1))
    obs, _ = genv.reset(options={"init_condition": ic, "beta": beta})
    total = torch.zeros(N, dtype=torch.float64, device=cuda_device)
    norm_sum = torch.zeros(N, dtype=torch.float64, device=cuda_device)
    for k in range(n_steps):
        obs, r, te, tr, _ = genv.step(ctl(obs))
        total += r
        norm_sum += obs.double().norm(dim=1)
    assert bool(te.all())
    total, norm_sum = total.cpu().numpy(), norm_sum.cpu().numpy()
    for i in range(N):
        u0 = (1, 10)[i % 2]
        assert abs(total[i] - want[u0][0]) <= RTOL * abs(want[u0][0]), (i, total[i], want[u0][0])
        assert abs(norm_sum[i] - want[u0][1]) <= 2 * RTOL * abs(want[u0][1]), (i, norm_sum[i], want[u0][1])

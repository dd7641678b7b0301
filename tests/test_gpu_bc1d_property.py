"""GPU property test (hypothesis): random grid sizes, substep counts, batch sizes, control / sensing / reward /
precision combinations of the Transport and ReactionDiffusion kernels against the numpy oracle.  States and
observations must be bit-identical in every drawn configuration."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import np_oracle as O
from tests.helpers import TOL_F32_REDUCTION, TOL_F64_REDUCTION, assert_rewards_close

pytestmark = pytest.mark.gpu

SUBSTEPS_TUNED = [1, 2, 4, 5, 10, 20, 25, 50, 100, 150, 300]


@st.composite
def configs(draw, large=False):
    env = draw(st.sampled_from(["rd", "tr"]))
    nx = draw(st.one_of(st.integers(511, 2047), st.integers(2048, 4300))) if large else draw(st.one_of(st.integers(4, 254), st.integers(255, 510)))
    reward = draw(st.sampled_from(["tuned", "temporal1", "temporal2", "temporalinf", "differential2"]))
    sub = draw(st.sampled_from(SUBSTEPS_TUNED + [3, 7, 30, 33, 64, 99, 101, 130])) if reward == "tuned" else draw(st.integers(1, 130))
    steps = draw(st.integers(2, 5))
    if reward == "tuned" and sub * steps <= 100:
        steps = 100 // sub + 2                 # the oracle needs nt > 100
    steps = min(steps, 120)
    control = draw(st.sampled_from(["Dirchilet", "Neumann"]))
    sensing = draw(st.sampled_from([("full", None), ("collocated", None), ("opposite", "Neumann")]))
    return dict(env=env, nx=nx, reward=reward, sub=sub, steps=steps, control=control, sensing=sensing,
                normalize=draw(st.booleans()), precision=draw(st.sampled_from(["ref", "fp64"])),
                beta64=draw(st.booleans()), act64=draw(st.booleans()), B=draw(st.integers(1, 9)),
                limit=draw(st.booleans()), seed=draw(st.integers(0, 2 ** 16)))


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow], derandomize=True)
@given(c=configs())
def test_random_configurations(cuda_device, c):
    _check(cuda_device, c)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow], derandomize=True)
@given(c=configs(large=True))
def test_random_configurations_large_grids(cuda_device, c):
    """The CTA-per-env kernels (more than 512 points per env): register-tiled and, for RD + Neumann, shared-memory."""
    _check(cuda_device, c)


def _check(cuda_device, c):
    import torch
    import pdecontrolgym_b200 as pcg
    env, nx, sub, steps, B = c["env"], c["nx"], c["sub"], c["steps"], c["B"]
    dx = 1.0 / nx
    if env == "rd":
        dt, npts = 0.4 * dx * dx, nx + 1
        if c["control"] == "Neumann" and c["normalize"]:
            steps = min(steps, 3)
    else:
        dt, npts = dx / 100, nx
    T = dt * sub * steps
    nt = int(round(T / dt) + 1)
    if nt - 1 != sub * steps or int(round(dt * sub / dt)) != sub or int(round(1 / dx)) != nx or nt <= 100:
        return                                     # floating-point round() in the ctor would change the shape: skip
    kw = dict(T=T, dt=dt, X=1, dx=dx, control_sample_rate=dt * sub, normalize=c["normalize"], sensing_loc=c["sensing"][0],
              control_type=c["control"], sensing_type=c["sensing"][1], limit_pde_state_size=c["limit"], max_state_value=25.0, max_control_value=20)
    nt_r = nt - 1
    rk = c["reward"]
    if rk == "tuned":
        ospec, prew = O.TunedReward1DSpec(nt_r, -1e3, 3e2), pcg.TunedReward1D(nt_r, -1e3, 3e2)
    else:
        hz = "differential" if rk.startswith("differential") else "temporal"
        name = rk[len(hz):]
        ospec = O.NormRewardSpec(nt_r, norm={"1": 1, "2": 2, "inf": np.inf}[name], horizon=hz)
        prew = pcg.NormReward(nt_r, norm=name, horizon=hz)
    rng = np.random.default_rng(c["seed"])
    ics = rng.uniform(0.5, 2.0, (B, 1)) * np.ones((B, npts)) + 0.1 * rng.standard_normal((B, npts))
    beta = (rng.uniform(-3, 3) * np.cos(rng.uniform(1, 9) * np.linspace(0, 1, npts))).astype(np.float64 if c["beta64"] else np.float32)
    Orc = O.ReactionDiffusionOracle if env == "rd" else O.TransportOracle
    oenv = Orc(B, reward=ospec, precision=c["precision"], **kw)
    eid = "PDEControlGym-ReactionDiffusionPDE1D" if env == "rd" else "PDEControlGym-TransportPDE1D"
    genv = pcg.make(eid, num_envs=B, device=cuda_device, precision=c["precision"], reward_class=prew, **kw)
    o_r = oenv.reset(ics, beta)
    o_g, _ = genv.reset(options={"init_condition": ics, "beta": beta})
    assert np.array_equal(o_g.cpu().numpy(), o_r)
    tol = TOL_F64_REDUCTION if c["precision"] == "fp64" else TOL_F32_REDUCTION
    for k in range(steps):
        a = (rng.uniform(-1, 1, B) if c["normalize"] else rng.uniform(-2, 2, B)).astype(np.float64 if c["act64"] else np.float32)
        o_r, r_r, te_r, tr_r = oenv.step(a)
        o_g, r_g, te_g, tr_g, _ = genv.step(torch.from_numpy(a))
        assert np.array_equal(o_g.cpu().numpy(), o_r, equal_nan=True), f"{c}: step {k}"
        assert np.array_equal(te_g.cpu().numpy(), te_r), f"{c}: terminated at step {k}"
        norm = np.sqrt((oenv.u.astype(np.float64) ** 2).sum(axis=1))
        near_limit = c["limit"] and np.any(np.abs(norm - 25.0) < 1e-3)
        if not near_limit:
            assert np.array_equal(tr_g.cpu().numpy(), tr_r), f"{c}: truncated at step {k}"
            scale = 2 * norm * (npts if rk.endswith("1") else 1) + np.abs(r_r)
            assert_rewards_close(r_g.cpu().numpy(), r_r, scale, tol, f"{c}: step {k}")

"""GPU parity at the FULL batch sizes of BASELINE.json (65,536 RD envs, 16,384 ARZ / BrainTumor envs, 1,024
Navier-Stokes envs, 2^20 Transport envs).  The oracle cannot run that many envs in seconds, so the batch is
built from R distinct envs tiled to the full size with per-env actions tiled the same way; then
  (1) every replica must be BIT-IDENTICAL to the first block (each of the N warps/CTAs did the same work), and
  (2) the first block must equal the oracle's result for those R envs (bit-identical states).
Together: every env of the full-size batch equals the oracle."""
import numpy as np
import pytest

from oracle import np_oracle as O
from tests.helpers import make_ics

pytestmark = pytest.mark.gpu


def _check_tiled(full, R):
    full = np.asarray(full)
    blocks = full.reshape(full.shape[0] // R, R, *full.shape[1:])
    assert np.array_equal(blocks, np.broadcast_to(blocks[0], blocks.shape), equal_nan=True), "replicas differ"
    return blocks[0]


@pytest.mark.parametrize("env,N,precision", [("rd", 65536, "fp64"), ("rd", 65536, "ref"), ("tr", 1 << 20, "ref")])
def test_bc1d_full_batch(cuda_device, env, N, precision):
    import torch
    import pdecontrolgym_b200 as pcg
    R, steps = 64, 3
    if env == "rd":
        kw = dict(T=1.0, dt=1e-5, X=1.0, dx=5e-3, control_sample_rate=1e-3)
        npts, nt_r = 201, 100000
        beta = 50 * np.cos(8 * np.arccos(np.linspace(0, 1, npts)))
        eid, Orc = "PDEControlGym-ReactionDiffusionPDE1D", O.ReactionDiffusionOracle
    else:
        kw = dict(T=5.0, dt=1e-4, X=1.0, dx=1e-2, control_sample_rate=0.1)
        npts, nt_r = 100, 50000
        beta = 5 * np.cos(7.35 * np.arccos(np.linspace(0, 1, npts)))
        eid, Orc = "PDEControlGym-TransportPDE1D", O.TransportOracle
        steps = 2
    beta = beta.astype(np.float64 if precision == "fp64" else np.float32)
    common = dict(normalize=True, max_control_value=20, **kw)
    rng = np.random.default_rng(0)
    ics = make_ics(rng, R, npts)
    genv = pcg.make(eid, num_envs=N, device=cuda_device, precision=precision, reward_class=pcg.NormReward(nt_r), **common)
    genv.reset(options={"init_condition": np.tile(ics, (N // R, 1)), "beta": beta})
    orc = Orc(R, reward=O.NormRewardSpec(nt_r, norm=2), precision="fp64" if precision == "fp64" else "ref", **common)
    orc.reset(ics, beta)
    for k in range(steps):
        a = rng.uniform(-1, 1, R).astype(np.float32)
        o_r, r_r, te_r, tr_r = orc.step(a)
        o_g, r_g, te_g, tr_g, _ = genv.step(torch.from_numpy(np.tile(a, N // R)))
        first = _check_tiled(o_g.cpu().numpy(), R)
        assert np.array_equal(first, o_r), f"step {k}: first block differs from the oracle"
        rew = _check_tiled(r_g.cpu().numpy(), R)
        np.testing.assert_allclose(rew, r_r, rtol=1e-12 if precision == "fp64" else 2e-6)
        assert not te_g.any() and not tr_g.any()


def test_arz_full_batch(cuda_device):
    import torch
    import pdecontrolgym_b200 as pcg
    N, R = 16384, 32
    kw = dict(T=240, dt=0.25, X=500, dx=10, simulation_type="outlet", control_freq=1)
    genv = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=cuda_device, reward_class=pcg.TrafficARZReward(), **kw)
    orc = O.TrafficARZOracle(R, **kw)
    genv.reset()
    orc.reset()
    rng = np.random.default_rng(1)
    for k in range(40):
        a = rng.uniform(0.96, 1.44, (R, 1))
        o_r, r_r, te_r, tr_r = orc.step(a)
        o_g, r_g, te_g, tr_g, _ = genv.step(torch.from_numpy(np.tile(a, (N // R, 1))))
        assert np.array_equal(_check_tiled(o_g.cpu().numpy(), R), o_r), f"step {k}"
        np.testing.assert_allclose(_check_tiled(r_g.cpu().numpy(), R), r_r, rtol=1e-12)


def test_brain_tumor_full_batch(cuda_device):
    import torch
    import pdecontrolgym_b200 as pcg
    N, R = 16384, 16
    P = dict(T=600, X=200, dt=1, dx=1, total_dosage=61.2)
    xs = np.linspace(0, 200, 201)
    ic = 0.8 * 1e5 * np.exp(-0.25 * xs ** 2)
    rng = np.random.default_rng(2)
    D, rho = 0.2 * rng.uniform(0.9, 1.1, R), 0.03 * rng.uniform(0.9, 1.1, R)
    genv = pcg.make("PDEControlGym-BrainTumor1D", num_envs=N, device=cuda_device, reward_class=pcg.BrainTumorReward(),
                    D=np.tile(D, N // R), rho=np.tile(rho, N // R), **P)
    orc = O.BrainTumorOracle(R, D=D, rho=rho, **P)
    genv.reset(options={"init_condition": ic})
    orc.reset(ic)
    for day in range(230):          # the whole Growth stage (bit-identical) and the first therapy days
        a = rng.uniform(0, 0.1, R)
        o_r, r_r, te_r, tr_r, v = orc.step(a)
        o_g, r_g, te_g, tr_g, info = genv.step(torch.from_numpy(np.tile(a, N // R)))
        first = _check_tiled(o_g.cpu().numpy(), R)
        if day < 200:
            assert np.array_equal(first, o_r), f"day {day}"
        else:
            np.testing.assert_allclose(first, o_r, rtol=1e-13)
        assert np.array_equal(_check_tiled(info["stage"].cpu().numpy(), R), orc.stage)


def test_navier_stokes_full_batch(cuda_device):
    import torch
    import pdecontrolgym_b200 as pcg
    N, R = 1024, 4
    bc = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"], "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}
    kw = dict(T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05, boundary_condition=bc, U_ref=np.zeros((200, 21, 21, 2)),
              action_ref=2.0 * np.ones(1000), maximum_pressure_iteration=2000)
    genv = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=cuda_device, reward_class=pcg.NSReward(0.1), action_dim=1, **kw)
    orc = O.NavierStokesOracle(R, gamma=0.1, **kw)
    rng = np.random.default_rng(3)
    c = rng.uniform(-5, 5, (3, R, 1, 1)) * np.ones((1, 21, 21))          # NS2Dppo.py:14-18 style constants
    genv.reset(options={"init": tuple(np.tile(c[i], (N // R, 1, 1)) for i in range(3))})
    orc.reset(c[0], c[1], c[2])
    for k in range(2):
        a = rng.uniform(2, 4, R)
        o_r, r_r, te_r, _ = orc.step(a)
        o_g, r_g, te_g, _, _ = genv.step(torch.from_numpy(np.tile(a, N // R)))
        assert np.array_equal(_check_tiled(o_g.cpu().numpy(), R), o_r), f"step {k}"
        np.testing.assert_allclose(_check_tiled(r_g.cpu().numpy(), R), r_r, rtol=1e-12)


@pytest.mark.parametrize("env,N,precision", [("rd", 65536, "fp64"), ("rd", 65536, "ref"), ("tr", 1 << 18, "ref")])
def test_bc1d_full_batch_all_distinct_vs_c_oracle(cuda_device, env, N, precision):
    """Every env of the full-size batch has its own initial condition, beta row and action; states are compared
    bit for bit with the plain-C oracle (oracle/c/pdecg_oracle.c, itself pinned to the numpy oracle and through
    it to the live reference)."""
    import torch
    import pdecontrolgym_b200 as pcg
    from oracle.c_oracle import Bc1dCOracle
    if env == "rd":
        kw = dict(T=1.0, dt=1e-5, X=1.0, dx=5e-3, control_sample_rate=1e-3)
        npts, nt_r, steps = 201, 100000, 2
        eid = "PDEControlGym-ReactionDiffusionPDE1D"
        base = 50 * np.cos(8 * np.arccos(np.linspace(0, 1, npts)))
    else:
        kw = dict(T=5.0, dt=1e-4, X=1.0, dx=1e-2, control_sample_rate=0.1)
        npts, nt_r, steps = 100, 50000, 1
        eid = "PDEControlGym-TransportPDE1D"
        base = 5 * np.cos(7.35 * np.arccos(np.linspace(0, 1, npts)))
    rng = np.random.default_rng(5)
    bdt = np.float64 if precision == "fp64" else np.float32
    ics = (rng.uniform(1, 10, (N, 1)) * np.ones((N, npts)) + 0.05 * rng.standard_normal((N, npts))).astype(np.float64)
    beta = (base[None, :] * rng.uniform(0.9, 1.1, (N, 1))).astype(bdt)
    common = dict(normalize=True, max_control_value=20, **kw)
    genv = pcg.make(eid, num_envs=N, device=cuda_device, precision=precision, reward_class=pcg.NormReward(nt_r), copy_obs=False, **common)
    corc = Bc1dCOracle(env, precision=precision, **common)
    genv.reset(options={"init_condition": ics, "beta": beta})
    corc.reset(ics, beta)
    for k in range(steps):
        a = rng.uniform(-1, 1, N).astype(np.float32)
        corc.step(a)
        o_g, r_g, te_g, tr_g, _ = genv.step(torch.from_numpy(a))
        got = o_g.cpu().numpy()
        assert np.array_equal(got, corc.u), f"step {k}: {np.count_nonzero((got != corc.u).any(axis=1))} of {N} envs differ"
        # NormReward (temporal, L2) from the oracle's rows
        want_r = -np.sqrt((corc.u.astype(np.float64) ** 2).sum(axis=1))
        np.testing.assert_allclose(r_g.cpu().numpy(), want_r, rtol=1e-12 if precision == "fp64" else 2e-6)
    assert (genv.time_index.cpu().numpy() == corc.t).all()

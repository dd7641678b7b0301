"""CUDA parity: NavierStokes2D kernel vs the numpy oracle (bit-identical fields: the scheme has no
transcendental and every point follows numpy's operation order) and vs the reference's golden
trajectory examples/NavierStokes/target.npz (K5), a 64 KB subset of which is committed under
tests/golden/."""
import os

import numpy as np
import pytest

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
BC_REF = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"],
          "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}
BC_MIX = {"upper": ["Controllable", "Neumann"], "lower": ["Neumann", "Dirchilet"],
          "left": ["Dirchilet", "Controllable"], "right": ["Neumann", "Neumann"]}


def make_pair(cuda_device, B, bc, iters, n=21, T=0.2, uref=None, gamma=0.1):
    import pdecontrolgym_b200 as pcg
    h = 1.0 / (n - 1)
    kw = dict(T=T, dt=1e-3 * (h / 0.05) ** 2, X=1, dx=h, Y=1, dy=h)
    nt = int(round(kw["T"] / kw["dt"]))
    if uref is None:
        uref = np.random.default_rng(11).uniform(-1, 1, (nt + 1, n, n, 2))
    aref = 2.0 * np.ones(1000)
    env = pcg.make("PDEControlGym-NavierStokes2D", num_envs=B, device=cuda_device, boundary_condition=bc, U_ref=uref,
                   action_ref=aref, maximum_pressure_iteration=iters, reward_class=pcg.NSReward(gamma), action_dim=1, **kw)
    orc = O.NavierStokesOracle(B, boundary_condition=bc, U_ref=uref, action_ref=aref, maximum_pressure_iteration=iters,
                               gamma=gamma, **kw)
    return env, orc, nt


@pytest.mark.parametrize("bc", [BC_REF, BC_MIX])
@pytest.mark.parametrize("iters", [0, 1, 2, 37, 150])
def test_vs_oracle(cuda_device, bc, iters):
    import torch
    B = 3
    env, orc, nt = make_pair(cuda_device, B, bc, iters)
    rng = np.random.default_rng(5)
    u0, v0, p0 = (rng.uniform(-1, 1, (B, 21, 21)) for _ in range(3))
    o_g, _ = env.reset(options={"init": (u0, v0, p0)})
    o_r = orc.reset(u0, v0, p0)
    assert np.array_equal(o_g.cpu().numpy(), o_r)
    for k in range(8):
        a = rng.uniform(2, 4, B)
        o_r, r_r, te_r, tr_r = orc.step(a)
        o_g, r_g, te_g, tr_g, info = env.step(torch.from_numpy(a))
        assert np.array_equal(o_g.cpu().numpy(), o_r), f"step {k}: fields not bit-identical (max |d| = {np.abs(o_g.cpu().numpy() - o_r).max()})"
        assert np.array_equal(env.p.cpu().numpy(), orc.p), f"step {k}: pressure"
        np.testing.assert_allclose(r_g.cpu().numpy(), r_r, rtol=1e-12)
        assert np.array_equal(te_g.cpu().numpy(), te_r) and not tr_g.any()


@pytest.mark.parametrize("n", [5, 8, 11, 16, 22, 23, 41, 45, 46, 64])      # 23..45: 512 threads x 4 points, 46..64: 1024 x 4
def test_grid_sizes(cuda_device, n):
    import torch
    env, orc, nt = make_pair(cuda_device, 2, BC_MIX, 25, n=n, T=0.02 * (20 / (n - 1)) ** 2)
    rng = np.random.default_rng(n)
    u0, v0, p0 = (rng.uniform(-1, 1, (2, n, n)) for _ in range(3))
    env.reset(options={"init": (u0, v0, p0)})
    orc.reset(u0, v0, p0)
    for k in range(min(6, nt - 1)):
        a = rng.uniform(-1, 1, 2)
        o_r, r_r, te_r, _ = orc.step(a)
        o_g, r_g, te_g, _, _ = env.step(torch.from_numpy(a))
        assert np.array_equal(o_g.cpu().numpy(), o_r)
        np.testing.assert_allclose(r_g.cpu().numpy(), r_r, rtol=1e-12)


def test_full_episode_terminates_and_guards(cuda_device):
    import torch
    env, orc, nt = make_pair(cuda_device, 2, BC_REF, 3, T=0.02)
    z = np.zeros((21, 21))
    env.reset(options={"init": (z, z, z)})
    orc.reset(z, z, z)
    for k in range(nt - 1):
        o_r, r_r, te_r, _ = orc.step(np.full(2, 3.0))
        o_g, r_g, te_g, _, info = env.step(np.full(2, 3.0))
        assert np.array_equal(o_g.cpu().numpy(), o_r) and np.array_equal(te_g.cpu().numpy(), te_r)
        assert info["valid"].all()
    assert te_r.all()
    o2, r2, te2, _, info = env.step(np.full(2, 3.0))       # the reference raises IndexError here
    assert not info["valid"].any() and te2.all() and np.array_equal(o2.cpu().numpy(), o_r)


def test_solve_pressure_entry_point(cuda_device):
    env, orc, nt = make_pair(cuda_device, 4, BC_REF, 60)
    rng = np.random.default_rng(2)
    u, v, p = (rng.uniform(-1, 1, (4, 21, 21)) for _ in range(3))
    got = env.solve_pressure(u, v, p).cpu().numpy()
    want = orc.solve_pressure(u, v, p)
    assert np.array_equal(got, want)


def test_k5_target_trajectory(cuda_device):
    """target.npz rows reproduced from u0 = target[0], v0 = p0 = 0, action(t) = 4 - 0.01 t, 2000 sweeps."""
    import torch
    g = np.load(os.path.join(GOLD, "ns_target_subset.npz"))
    rows, u_rows, v_rows, u0 = g["rows"], g["u"], g["v"], g["u0"]
    uref = np.zeros((200, 21, 21, 2))
    env, orc, nt = make_pair(cuda_device, 2, BC_REF, 2000, uref=uref)
    z = np.zeros((21, 21))
    env.reset(options={"init": (u0, z, z)})
    want = {int(r): (u_rows[i], v_rows[i]) for i, r in enumerate(rows)}
    for t in range(1, 200):
        obs, r, te, tr, _ = env.step(np.full(2, 4 - 0.01 * t))
        if t in want:
            o = obs.cpu().numpy()[1]
            np.testing.assert_allclose(o[:, :, 0], want[t][0], rtol=0, atol=1e-12)
            np.testing.assert_allclose(o[:, :, 1], want[t][1], rtol=0, atol=1e-12)
    assert te.all()


def test_grid_beyond_shared_memory_is_rejected_loudly(cuda_device):
    import pdecontrolgym_b200 as pcg
    env, orc, nt = make_pair(cuda_device, 1, BC_REF, 2, n=65, T=0.002)
    z = np.zeros((65, 65))
    with pytest.raises(pcg.PdecgError, match="exceeds the shared-memory-resident limit"):
        env.reset(options={"init": (z, z, z)})

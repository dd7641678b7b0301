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


def make_pair(cuda_device, B, bc, iters, n=21, T=0.2, uref=None, gamma=0.1, **env_kw):
    import pdecontrolgym_b200 as pcg
    h = 1.0 / (n - 1)
    kw = dict(T=T, dt=1e-3 * (h / 0.05) ** 2, X=1, dx=h, Y=1, dy=h)
    nt = int(round(kw["T"] / kw["dt"]))
    if uref is None:
        uref = np.random.default_rng(11).uniform(-1, 1, (nt + 1, n, n, 2))
    aref = 2.0 * np.ones(1000)
    env = pcg.make("PDEControlGym-NavierStokes2D", num_envs=B, device=cuda_device, boundary_condition=bc, U_ref=uref,
                   action_ref=aref, maximum_pressure_iteration=iters, reward_class=pcg.NSReward(gamma), action_dim=1, **kw, **env_kw)
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


@pytest.mark.parametrize("jacobi", ["registers", "shared_memory"])
@pytest.mark.parametrize("n", [5, 7, 8, 11, 12, 16, 18, 22, 23, 34, 41, 45, 46, 50, 63, 64])
def test_grid_sizes(cuda_device, monkeypatch, n, jacobi):
    """On-chip grids other than 21 x 21: the CTA kernel with the register-blocked Jacobi (16 thread columns x 1..5 points,
    5-point thread rows; the sizes straddle its block / warp boundaries) and, behind PDECG_NS_SMEM_JACOBI, the
    shared-memory Jacobi (256 x 2, 512 x 4, 1024 x 4 threads x points)."""
    import torch
    if jacobi == "shared_memory":
        if n in (7, 12, 18, 34, 50, 63):
            pytest.skip("one size per instantiation is enough for the fallback form")
        monkeypatch.setenv("PDECG_NS_SMEM_JACOBI", "1")
    else:
        monkeypatch.setenv("PDECG_NS_BLK_JACOBI", "1")          # also beyond 50 x 50, where it is not the default
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


def _run_pair(env, orc, rng, B, n, steps):
    import torch
    u0, v0, p0 = (rng.uniform(-1, 1, (B, n, n)) for _ in range(3))
    o_g, _ = env.reset(options={"init": (u0, v0, p0)})
    o_r = orc.reset(u0, v0, p0)
    assert np.array_equal(o_g.cpu().numpy(), o_r)
    for k in range(steps):
        a = rng.uniform(2, 4, B)
        o_r, r_r, te_r, _ = orc.step(a)
        o_g, r_g, te_g, _, info = env.step(torch.from_numpy(a))
        assert np.array_equal(env.p.cpu().numpy(), orc.p), f"step {k}: pressure not bit-identical (max |d| = {np.abs(env.p.cpu().numpy() - orc.p).max()})"
        assert np.array_equal(o_g.cpu().numpy(), o_r), f"step {k}: fields not bit-identical"
        np.testing.assert_allclose(r_g.cpu().numpy(), r_r, rtol=1e-12)
        assert np.array_equal(te_g.cpu().numpy(), te_r)


@pytest.mark.parametrize("n,iters", [(21, 0), (21, 1), (21, 2), (21, 37), (22, 23), (45, 11), (62, 20), (63, 19), (64, 31), (65, 12), (70, 21), (71, 25), (83, 30), (131, 52)])
def test_tiled_hbm_path_on_small_and_medium_grids(cuda_device, monkeypatch, n, iters):
    """The HBM-resident path (TMA-tiled, temporally blocked Jacobi + whole-grid stencil kernels) forced onto grids the
    on-chip kernels also serve, and just beyond them: u, v, p bit-identical to the oracle, sweep counts that are not
    multiples of the temporal depth (10), odd and even widths (padded pitch), one tile and several tiles per env, with the
    domain's bottom / right boundary at every slot of a thread's 5 x 5 block (n mod 5 = 0 ... 4)."""
    monkeypatch.setenv("PDECG_NS_TILED", "1")
    B = 3
    env, orc, nt = make_pair(cuda_device, B, BC_MIX, iters, n=n, T=0.2 * (20 / (n - 1)) ** 2)
    _run_pair(env, orc, np.random.default_rng(n + iters), B, n, 4)


@pytest.mark.parametrize("n", [128, 200])
def test_large_grids_vs_oracle(cuda_device, n):
    """Grids beyond the shared-memory-resident limit (the reference never runs them: 2000 numpy sweeps of 128 x 128 take
    seconds per step): 3 steps of 60 sweeps, bit-identical u, v, p; the oracle is pinned to the live reference at 128 x 128
    in tests/test_oracle_vs_reference.py::test_navier_stokes."""
    B = 2
    env, orc, nt = make_pair(cuda_device, B, BC_REF, 60, n=n, T=0.2 * (20 / (n - 1)) ** 2)
    _run_pair(env, orc, np.random.default_rng(n), B, n, 3)


def test_large_grid_full_solve_properties(cuda_device):
    """256 x 256 x 2000 sweeps (the oracle needs minutes for it): size-independent properties instead -- (1) envs with the
    same inputs agree bit for bit whatever their position in the batch (different SMs, different tile schedules), (2) the
    tiled solve equals repeated applications of the one-sweep path (`maximum_pressure_iteration=1`, which never enters the
    tiled kernel), (3) the pressure ring equals its nearest interior value."""
    import torch
    n, B = 256, 5
    rng = np.random.default_rng(3)
    u0 = rng.uniform(-1, 1, (1, n, n)).repeat(B, 0)
    v0 = rng.uniform(-1, 1, (1, n, n)).repeat(B, 0)
    p0 = rng.uniform(-1, 1, (1, n, n)).repeat(B, 0)
    env, orc, nt = make_pair(cuda_device, B, BC_REF, 2000, n=n, T=0.2 * (20 / (n - 1)) ** 2)
    env.reset(options={"init": (u0, v0, p0)})
    a = torch.full((B,), 3.0, dtype=torch.float64)
    env.step(a)
    u, v, p = env.u.cpu().numpy(), env.v.cpu().numpy(), env.p.cpu().numpy()
    for f in (u, v, p):
        assert np.isfinite(f).all() and all(np.array_equal(f[0], f[b]) for b in range(1, B))
    assert np.array_equal(p[:, 0, :], p[:, 1, :]) and np.array_equal(p[:, -1, :], p[:, -2, :])
    assert np.array_equal(p[:, 1:-1, 0], p[:, 1:-1, 1]) and np.array_equal(p[:, 1:-1, -1], p[:, 1:-1, -2])
    # (2) 60 sweeps of the same right-hand side: tiled (6 launches of 10 after the first sweep) vs sweep by sweep
    env60, _, _ = make_pair(cuda_device, 1, BC_REF, 60, n=n, T=0.2 * (20 / (n - 1)) ** 2)
    env1, _, _ = make_pair(cuda_device, 1, BC_REF, 1, n=n, T=0.2 * (20 / (n - 1)) ** 2)
    us, vs = rng.uniform(-1, 1, (1, n, n)), rng.uniform(-1, 1, (1, n, n))
    want = torch.from_numpy(p0[:1].copy()).to(cuda_device)
    for _ in range(60):
        want = env1.solve_pressure(us, vs, want)
    got = env60.solve_pressure(us, vs, p0[:1])
    assert torch.equal(got, want)


@pytest.mark.parametrize("lpe32", [False, True])
@pytest.mark.parametrize("B", [1, 2, 5, 8])
def test_warp_kernel_mappings_odd_batches_and_stale_partners(cuda_device, monkeypatch, B, lpe32):
    """21 x 21 register kernel, both lane mappings (two envs per warp: 4 x 4 lanes of 5 x 5 points; one env per warp:
    10 x 3 lanes of 2 x 7): odd batches leave half a warp without an env, and envs at different time indices put a live env
    next to one that is past its last step -- fields stay bit-identical to the oracle for every env, 2000 sweeps included."""
    import torch
    if lpe32:
        monkeypatch.setenv("PDECG_NS_LPE32", "1")
    for iters in (5, 2000):
        env, orc, nt = make_pair(cuda_device, B, BC_MIX, iters, T=0.004)      # 3 steps per episode
        rng = np.random.default_rng(100 + B)
        u0, v0, p0 = (rng.uniform(-1, 1, (B, 21, 21)) for _ in range(3))
        env.reset(options={"init": (u0, v0, p0)})
        orc.reset(u0, v0, p0)
        # stagger: every other env starts one step late (its partner in the warp finishes earlier)
        late = np.arange(B) % 2 == 1
        a0 = rng.uniform(2, 4, B)
        env.step(torch.from_numpy(a0))
        orc.step(a0)
        if late.any():
            env.reset(options={"init": (u0, v0, p0), "mask": late})
            orc_late = O.NavierStokesOracle(B, boundary_condition=BC_MIX, U_ref=orc.U_ref, action_ref=orc.action_ref,
                                            maximum_pressure_iteration=iters, gamma=0.1, T=0.004, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05)
            orc_late.reset(u0, v0, p0)
        for k in range(3):
            a = rng.uniform(2, 4, B)
            o_g, r_g, te_g, _, info = env.step(torch.from_numpy(a))
            og = o_g.cpu().numpy()
            valid = info["valid"].cpu().numpy()
            if k < 2:                        # early envs: steps 2..3 of their episode (the 3rd call is past the end)
                o_e, r_e, te_e, _ = orc.step(a)
                assert np.array_equal(og[~late], o_e[~late]) and np.array_equal(env.p.cpu().numpy()[~late], orc.p[~late]), f"iters {iters} step {k}"
                np.testing.assert_allclose(r_g.cpu().numpy()[~late], r_e[~late], rtol=1e-12)
                assert valid[~late].all()
            else:
                assert not valid[~late].any() and np.array_equal(og[~late], o_e[~late])
            if late.any():                   # late envs: steps 1..3 of theirs
                o_l, r_l, te_l, _ = orc_late.step(a)
                assert np.array_equal(og[late], o_l[late]) and np.array_equal(env.p.cpu().numpy()[late], orc_late.p[late]), f"iters {iters} step {k} (late)"
                np.testing.assert_allclose(r_g.cpu().numpy()[late], r_l[late], rtol=1e-12)
                assert valid[late].all() and np.array_equal(te_g.cpu().numpy()[late], te_l[late])


@pytest.mark.parametrize("n,iters,mapping", [(21, 2000, "two_envs_per_warp"), (21, 300, "one_env_per_warp"), (33, 400, "cta"), (64, 120, "cta")])
def test_fp32_fast_mode_tolerance(cuda_device, monkeypatch, n, iters, mapping):
    """precision="fp32" (opt-in): float32 fields / observation / arithmetic in the reference's operation order; actions,
    U_ref, the reward and the initial fields stay float64.  Stated tolerance against the float64 oracle over a whole
    episode from a smooth initial state with the reference's boundary set: fields <= 2e-4 of the field scale per step
    (Jacobi is a contraction: rounding errors of earlier sweeps are damped, not accumulated), pressure <= 2e-3 of its
    scale, reward <= 1e-4 relative; termination flags identical.  The float32 state is returned as float32."""
    import torch
    if mapping == "one_env_per_warp":
        monkeypatch.setenv("PDECG_NS_LPE32", "1")
    B = 3
    T = 0.03 * (20 / (n - 1)) ** 2
    env, orc, nt = make_pair(cuda_device, B, BC_REF, iters, n=n, T=T, precision="fp32")
    xs = np.linspace(0, 1, n)
    X, Y = np.meshgrid(xs, xs)
    u0 = np.stack([0.3 * (b + 1) * np.sin(np.pi * X) * np.cos(np.pi * Y) for b in range(B)])
    v0 = np.stack([-0.2 * (b + 1) * np.cos(np.pi * X) * np.sin(np.pi * Y) for b in range(B)])
    p0 = np.zeros_like(u0)
    o_g, _ = env.reset(options={"init": (u0, v0, p0)})
    o_r = orc.reset(u0, v0, p0)
    assert o_g.dtype == torch.float32 and env.u.dtype == torch.float32 and env.p.dtype == torch.float32
    assert np.array_equal(o_g.cpu().numpy(), o_r.astype(np.float32))
    rng = np.random.default_rng(3)
    steps = 0
    while True:
        a = rng.uniform(1, 3, B)
        o_r, r_r, te_r, tr_r = orc.step(a)
        o_g, r_g, te_g, tr_g, info = env.step(torch.from_numpy(a))
        steps += 1
        scale = np.abs(o_r).max()
        assert np.abs(o_g.cpu().numpy() - o_r).max() <= 2e-4 * scale, f"step {steps}: fields off by {np.abs(o_g.cpu().numpy() - o_r).max() / scale:.2e} of scale"
        assert np.abs(env.p.cpu().numpy() - orc.p).max() <= 2e-3 * max(np.abs(orc.p).max(), 1e-12), f"step {steps}: pressure"
        np.testing.assert_allclose(r_g.cpu().numpy(), r_r, rtol=1e-4)
        assert r_g.dtype == torch.float64
        assert np.array_equal(te_g.cpu().numpy(), te_r) and not tr_g.any()
        if te_r.all():
            break
    assert steps == nt - 1


def test_fp32_rejects_what_it_does_not_cover(cuda_device):
    import pdecontrolgym_b200 as pcg
    env, _, _ = make_pair(cuda_device, 1, BC_REF, 5, precision="fp32")
    z = np.zeros((21, 21))
    env.reset(options={"init": (z, z, z)})
    with pytest.raises(NotImplementedError):
        env.solve_pressure(z[None], z[None], z[None])
    with pytest.raises(Exception, match="on-chip"):
        e2, _, _ = make_pair(cuda_device, 1, BC_REF, 5, n=81, T=0.002, precision="fp32")
        e2.reset(options={"init": (np.zeros((81, 81)),) * 3})


@pytest.mark.parametrize("iters", [0, 1, 2, 3, 38])
@pytest.mark.parametrize("n", [30, 52])
def test_register_blocked_jacobi_sweep_counts(cuda_device, monkeypatch, n, iters):
    """Odd / even / degenerate sweep counts of the register-blocked Jacobi (first sweep from shared memory, register
    ping-pong, tail sweep), warm-started from a random pressure field with non-trivial boundary values."""
    import torch
    monkeypatch.setenv("PDECG_NS_BLK_JACOBI", "1")
    env, orc, nt = make_pair(cuda_device, 3, BC_MIX, iters, n=n, T=0.02 * (20 / (n - 1)) ** 2)
    rng = np.random.default_rng(100 * n + iters)
    u0, v0, p0 = (rng.uniform(-1, 1, (3, n, n)) for _ in range(3))
    env.reset(options={"init": (u0, v0, p0)})
    orc.reset(u0, v0, p0)
    for k in range(min(4, nt - 1)):
        a = rng.uniform(-2, 2, 3)
        o_r, r_r, te_r, _ = orc.step(a)
        o_g, r_g, te_g, _, _ = env.step(torch.from_numpy(a))
        assert np.array_equal(o_g.cpu().numpy(), o_r), f"step {k}: fields"
        assert np.array_equal(env.p.cpu().numpy(), orc.p), f"step {k}: pressure"
        np.testing.assert_allclose(r_g.cpu().numpy(), r_r, rtol=1e-12)


def test_tiled_solve_equals_sweep_by_sweep_at_512(cuda_device):
    """512 x 512 (81 tiles per env): 40 sweeps through the tiled kernel (4 launches of 10 after the first sweep) equal 40
    applications of the one-sweep path, bit for bit, for two envs with different fields."""
    import torch
    n, B = 512, 2
    rng = np.random.default_rng(512)
    kw = dict(n=n, T=0.2 * (20 / (n - 1)) ** 2)
    env40, _, _ = make_pair(cuda_device, B, BC_REF, 40, **kw)
    env1, _, _ = make_pair(cuda_device, B, BC_REF, 1, **kw)
    us, vs, p0 = (rng.uniform(-1, 1, (B, n, n)) for _ in range(3))
    want = torch.from_numpy(p0.copy()).to(cuda_device)
    for _ in range(40):
        want = env1.solve_pressure(us, vs, want)
    got = env40.solve_pressure(us, vs, p0)
    assert torch.equal(got, want)
    p = got.cpu().numpy()
    assert np.array_equal(p[:, 0, :], p[:, 1, :]) and np.array_equal(p[:, 1:-1, -1], p[:, 1:-1, -2])

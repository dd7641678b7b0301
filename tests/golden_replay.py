"""Replays the committed reference runs (tests/golden/*.npz, produced by tests/golden/gen_golden.py from the
unmodified reference) through either the numpy oracle (CPU tests) or the CUDA envs (GPU tests)."""
import os

import numpy as np

from oracle import np_oracle as O
from tests.golden.gen_golden import BC1D_CASES, BC1D_LARGE_CASES, BT_P, bt_ic
from tests.helpers import bc1d_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def _to_np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


# ---- Transport / ReactionDiffusion --------------------------------------------------------------------
def bc1d_cases():
    return [c[0] for c in BC1D_CASES] + [c[0] for c in BC1D_LARGE_CASES]


def replay_bc1d(name, backend, device=None, reward_rtol=0.0):
    """backend: 'oracle' or 'cuda'.  States/observations must equal the reference bit for bit; rewards
    exactly for the oracle (same numpy routines), within `reward_rtol` of the row norms for CUDA."""
    case = next(c for c in BC1D_CASES + BC1D_LARGE_CASES if c[0] == name)
    (_, env, ct, sl, st, nz, bd, ad, rk, f64) = case[:10]
    nx, steps = (case[10], case[11]) if len(case) > 10 else (None, 12)       # large-grid cases carry their own size
    g = load("bc1d_large_grid_reference_runs.npz" if len(case) > 10 else "bc1d_reference_runs.npz")
    kw, npts, beta, nt = bc1d_case(env, nx, steps)
    ic, beta, acts = g[name + "/ic"], g[name + "/beta"], g[name + "/actions"]
    common = dict(normalize=nz, sensing_loc=sl, control_type=ct, sensing_type=st, limit_pde_state_size=False,
                  max_state_value=30.0, max_control_value=20, **kw)
    nt_r = nt - 1
    if backend == "oracle":
        spec = O.TunedReward1DSpec(nt_r, -1e3, 3e2) if rk == "tuned" else O.NormRewardSpec(nt_r, norm=rk[1], horizon=rk[0])
        Orc = O.ReactionDiffusionOracle if env == "rd" else O.TransportOracle
        e = Orc(1, reward=spec, precision="fp64" if f64 else "ref", **common)
        obs = e.reset(ic, beta)
        step = lambda a: e.step(a)[:4]  # noqa: E731
    else:
        import pdecontrolgym_b200 as pcg
        rw = pcg.TunedReward1D(nt_r, -1e3, 3e2) if rk == "tuned" else pcg.NormReward(
            nt_r, norm={1: "1", 2: "2", np.inf: "inf"}[rk[1]], horizon=rk[0])
        eid = "PDEControlGym-ReactionDiffusionPDE1D" if env == "rd" else "PDEControlGym-TransportPDE1D"
        e = pcg.make(eid, num_envs=1, device=device, precision="fp64" if f64 else "ref", reward_class=rw, **common)
        obs, _ = e.reset(options={"init_condition": ic, "beta": beta})
        step = lambda a: e.step(a)[:4]  # noqa: E731
    want = g[name + "/obs"]
    assert np.array_equal(_to_np(obs).reshape(want[0].shape), want[0]), f"{name}: reset obs"
    for k in range(steps):
        o, r, te, tr = step(np.array([acts[k]], dtype=acts.dtype))
        o = _to_np(o).reshape(want[k + 1].shape)
        assert o.dtype == want.dtype, (o.dtype, want.dtype)
        assert np.array_equal(o, want[k + 1], equal_nan=True), f"{name} step {k}: obs differs from the reference run"
        r = float(_to_np(r).reshape(-1)[0])
        rw_ref = float(g[name + "/reward"][k])
        if backend == "oracle" or reward_rtol == 0.0:
            assert r == rw_ref, f"{name} step {k}: reward {r!r} != {rw_ref!r}"
        else:
            scale = 2 * np.sqrt((want[k + 1].astype(np.float64) ** 2).sum() * (npts if o.size == 1 else 1)) + abs(rw_ref) + 1.0
            assert abs(r - rw_ref) <= reward_rtol * scale, f"{name} step {k}: reward {r} vs {rw_ref}"
        assert bool(_to_np(te).reshape(-1)[0]) == bool(g[name + "/terminated"][k])
        assert bool(_to_np(tr).reshape(-1)[0]) == bool(g[name + "/truncated"][k])


# ---- Traffic ARZ -----------------------------------------------------------------------------------------
def arz_cases():
    return [f"{st}_cf{cf}" for st in ("inlet", "outlet", "both") for cf in (1, 2)]


def replay_arz(key, backend, device=None):
    g = load("arz_reference_runs.npz")
    st, cf = key.rsplit("_cf", 1)
    kw = dict(T=5, dt=0.25, X=500, dx=10, simulation_type=st, control_freq=int(cf), limit_pde_state_size=True)
    if backend == "oracle":
        e = O.TrafficARZOracle(1, **kw)
        obs = e.reset()
    else:
        import pdecontrolgym_b200 as pcg
        e = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=1, device=device, reward_class=pcg.TrafficARZReward(), **kw)
        obs, _ = e.reset()
    want = g[key + "/obs"]
    assert np.array_equal(_to_np(obs)[0], want[0])
    acts = g[key + "/actions"]
    for k in range(100):
        o, r, te, tr = e.step(acts[k][None, :])[:4]
        if (k + 1) % 5 == 0:
            assert np.array_equal(_to_np(o)[0], want[(k + 1) // 5]), f"{key} step {k}"
        np.testing.assert_allclose(float(_to_np(r)[0]), g[key + "/reward"][k], rtol=0 if backend == "oracle" else 1e-12)
        assert bool(_to_np(te)[0]) == bool(g[key + "/terminated"][k]) and bool(_to_np(tr)[0]) == bool(g[key + "/truncated"][k])


# ---- BrainTumor ------------------------------------------------------------------------------------------
def replay_brain_tumor_raw(backend, device=None):
    g = load("brain_tumor_reference_runs.npz")
    days = {int(d): i for i, d in enumerate(g["raw/obs_days"])}
    n = len(g["raw/reward"])
    if backend == "oracle":
        e = O.BrainTumorOracle(1, **BT_P)
        e.reset(bt_ic(200, 201))
        e.t_benchmark = np.array([363])
        step = lambda: e.step(np.array([0.05]))  # noqa: E731
        stage_of = lambda out: int(e.stage[0])  # noqa: E731
    else:
        import pdecontrolgym_b200 as pcg
        e = pcg.make("PDEControlGym-BrainTumor1D", num_envs=1, device=device, reward_class=pcg.BrainTumorReward(),
                     reset_init_condition_func=bt_ic, **BT_P)
        e.reset()
        e.t_benchmark = [363]
        step = lambda: e.step(np.array([0.05]))  # noqa: E731
        stage_of = lambda out: int(out[4]["stage"][0])  # noqa: E731
    rtol = 0.0 if backend == "oracle" else 1e-13
    for d in range(1, n + 1):
        out = step()
        o, r, te, tr = out[:4]
        if d in days:
            np.testing.assert_allclose(_to_np(o)[0], g["raw/obs"][days[d]], rtol=rtol, atol=0, err_msg=f"day {d}")
        np.testing.assert_allclose(float(_to_np(r)[0]), g["raw/reward"][d - 1], rtol=0 if backend == "oracle" else 1e-10, atol=1e-12)
        assert bool(_to_np(te)[0]) == bool(g["raw/terminated"][d - 1]) and bool(_to_np(tr)[0]) == bool(g["raw/truncated"][d - 1])
        assert stage_of(out) == int(g["raw/stage"][d - 1]), f"day {d}: stage"
    assert bool(_to_np(te)[0]) or bool(_to_np(tr)[0])


def replay_brain_tumor_wrapper(name, backend, device=None):
    g = load("brain_tumor_reference_runs.npz")
    tb, growth, first_therapy, first_post, death, n_outer = (int(x) for x in g[f"wrap_{name}/summary"])
    props = [] if name == "open" else [1 / 34] * 34
    weekends = name.endswith("weekends")
    if backend == "oracle":
        env = O.BrainTumorOracle(1, **BT_P)
        w = O.TherapyWrapperOracle(env, bt_ic(200, 201), weekends=weekends)
        assert w.benchmark() == tb
        o = w.reset()
        get = lambda a: int(np.asarray(a).reshape(-1)[0])  # noqa: E731
    else:
        import pdecontrolgym_b200 as pcg
        env = pcg.make("PDEControlGym-BrainTumor1D", num_envs=1, device=device, reward_class=pcg.BrainTumorReward(),
                       reset_init_condition_func=bt_ic, **BT_P)
        w = pcg.TherapyWrapper(env, weekends=weekends)
        assert int(w.benchmark()[0]) == tb
        o, _ = w.reset()
        get = lambda a: int(_to_np(a).reshape(-1)[0])  # noqa: E731
    rtol = 0.0 if backend == "oracle" else 1e-13
    np.testing.assert_allclose(_to_np(o)[0], g[f"wrap_{name}/obs_first_last"][0], rtol=rtol)
    assert get(env.growthDays) == growth and get(env.firstTherapyDay) == first_therapy
    rewards = g[f"wrap_{name}/reward"]
    for k in range(n_outer):
        a = props[k] if k < len(props) else 0.0
        o, r, te, tr = w.step(a if backend == "oracle" else np.array([a]))[:4]
        np.testing.assert_allclose(float(_to_np(r).reshape(-1)[0]), rewards[k], rtol=0 if backend == "oracle" else 1e-10, atol=1e-12)
        if k == 0:
            np.testing.assert_allclose(_to_np(o)[0], g[f"wrap_{name}/obs_first_last"][1], rtol=rtol)
    assert bool(_to_np(te).reshape(-1)[0]) or bool(_to_np(tr).reshape(-1)[0])
    np.testing.assert_allclose(_to_np(o)[0], g[f"wrap_{name}/obs_first_last"][2], rtol=rtol)
    assert get(env.cDeathDay) == death and get(env.firstPostTherapyDay) == first_post


# ---- Navier-Stokes ---------------------------------------------------------------------------------------
def replay_ns(backend, device=None):
    g = load("ns_reference_run.npz")
    bc = {"upper": ["Controllable", "Neumann"], "lower": ["Neumann", "Dirchilet"], "left": ["Dirchilet", "Controllable"], "right": ["Neumann", "Neumann"]}
    kw = dict(T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05, boundary_condition=bc, U_ref=g["uref_rows"],
              action_ref=2.0 * np.ones(1000), maximum_pressure_iteration=37)
    if backend == "oracle":
        e = O.NavierStokesOracle(1, gamma=0.1, **kw)
        e.reset(g["u0"], g["v0"], g["p0"])
        pressure = lambda: e.p[0]  # noqa: E731
    else:
        import pdecontrolgym_b200 as pcg
        e = pcg.make("PDEControlGym-NavierStokes2D", num_envs=1, device=device, reward_class=pcg.NSReward(0.1), action_dim=1, **kw)
        e.reset(options={"init": (g["u0"], g["v0"], g["p0"])})
        pressure = lambda: _to_np(e.p)[0]  # noqa: E731
    for k in range(6):
        o, r, te, tr = e.step(np.array([g["actions"][k]]))[:4]
        assert np.array_equal(_to_np(o)[0], g["obs"][k]), f"step {k}: fields differ from the reference run"
        assert np.array_equal(pressure(), g["p"][k]), f"step {k}: pressure"
        np.testing.assert_allclose(float(_to_np(r)[0]), g["reward"][k], rtol=0 if backend == "oracle" else 1e-12)

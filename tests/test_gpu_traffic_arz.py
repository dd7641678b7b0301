"""CUDA parity: TrafficPDE1D kernel vs the numpy oracle.  States / observations BIT-IDENTICAL (no
transcendental in the step), rewards (two L2 norms) to 1e-12 relative, flags exact.  The reference
publishes no golden for this env; pin S1 below was produced by running the reference."""
import numpy as np
import pytest

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def run(cuda_device, st, cf, limit, steps=400, B=4, T=20, X=500, dx=10, dt=0.25, rs=None, seed=3, rs_init=None):
    import torch
    import pdecontrolgym_b200 as pcg
    kw = dict(T=T, dt=dt, X=X, dx=dx)
    if rs_init is not None:
        kw["rs_init"] = rs_init          # 'outlet-train': the constructor's draw, which fixes the action bounds for good
    env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=B, device=cuda_device, simulation_type=st, control_freq=cf,
                   limit_pde_state_size=limit, reward_class=pcg.TrafficARZReward(), **kw)
    orc = O.TrafficARZOracle(B, simulation_type=st, control_freq=cf, limit_pde_state_size=limit, **kw)
    if st == "outlet-train":
        o_g, _ = env.reset(options={"rs": rs})
        o_r = orc.reset(rs=rs)
    else:
        o_g, _ = env.reset()
        o_r = orc.reset()
    assert np.array_equal(o_g.cpu().numpy(), o_r)
    rng = np.random.default_rng(seed)
    nact = 2 if st == "both" else 1
    n_term = 0
    for k in range(steps):
        a = rng.uniform(0.7 * 1.2, 1.3 * 1.2, size=(B, nact))
        o_r, r_r, te_r, tr_r = orc.step(a)
        o_g, r_g, te_g, tr_g, info = env.step(torch.from_numpy(a))
        assert np.array_equal(o_g.cpu().numpy(), o_r), f"step {k}: obs not bit-identical"
        np.testing.assert_allclose(r_g.cpu().numpy(), r_r, rtol=1e-12, err_msg=f"step {k}")
        assert np.array_equal(te_g.cpu().numpy(), te_r) and np.array_equal(tr_g.cpu().numpy(), tr_r), f"step {k}: flags"
        assert np.array_equal(env.time_index.cpu().numpy(), orc.time_index)
        n_term += int(te_r.any())
    return n_term


@pytest.mark.parametrize("st", ["inlet", "outlet", "both"])
@pytest.mark.parametrize("cf", [1, 3])
@pytest.mark.parametrize("limit", [False, True])
def test_vs_oracle(cuda_device, st, cf, limit):
    assert run(cuda_device, st, cf, limit) >= 1       # crosses the time-based termination + time reset


def test_outlet_train(cuda_device):
    """reset() re-draws rs / qs, but actions keep being clipped to the bounds built from the CONSTRUCTOR's draw
    (traffic_arz_env.py:97-100 vs :251-256): every env here resets to an rs different from its construction-time one."""
    run(cuda_device, "outlet-train", 2, False, rs=np.array([0.115, 0.12, 0.125, 0.12]), rs_init=np.array([0.125, 0.115, 0.12, 0.125]))
    run(cuda_device, "outlet-train", 1, True, rs=np.array([0.12, 0.125, 0.115, 0.115]), rs_init=np.array([0.12, 0.12, 0.12, 0.12]), steps=120)


@pytest.mark.parametrize("X,dx", [(100, 10), (310, 10), (640, 10), (500, 4), (500, 2), (500, 1.75), (500, 1.4), (500, 1.2), (500, 1)])   # ... 251, 287, 358, 418, 501 points
def test_grid_sizes(cuda_device, X, dx):
    run(cuda_device, "both", 2, True, steps=60, B=3, X=X, dx=dx, dt=0.25 * dx / 10)


def test_pin_s1_full_episode(cuda_device):
    """SURVEY.md Appendix B pin S1: tutorial defaults, action == qs: 3,840 steps, first / last / sum reward."""
    import torch
    import pdecontrolgym_b200 as pcg
    env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=2, device=cuda_device, simulation_type="outlet",
                   reward_class=pcg.TrafficARZReward(), T=240, dt=0.25, X=500, dx=10, copy_obs=False)
    env.reset()
    a = torch.full((2, 1), 1.2, dtype=torch.float64, device=cuda_device)
    rewards, n = [], 0
    while True:
        _, r, te, tr, _ = env.step(a)
        rewards.append(r)
        n += 1
        if n % 256 == 0 or n >= 3840:
            if bool(te.any()) or bool(tr.any()) or n > 4000:
                break
    rew = torch.stack(rewards).cpu().numpy()
    assert n == 3840
    np.testing.assert_allclose(rew[0], -0.9956612448005211, rtol=1e-12)
    np.testing.assert_allclose(rew[-1], -0.2512486635597482, rtol=1e-12)
    np.testing.assert_allclose(rew.sum(axis=0), -1106.0954718055284, rtol=1e-11)


def test_errors(cuda_device):
    import pdecontrolgym_b200 as pcg
    kw = dict(T=20, dt=0.25, X=500, dx=10, num_envs=1, device=cuda_device)
    with pytest.raises(ValueError, match="Invalid simulation type"):
        pcg.make("PDEControlGym-TrafficPDE1D", simulation_type="sideways", **kw)
    with pytest.raises(ValueError, match="equilibrium"):
        pcg.make("PDEControlGym-TrafficPDE1D", v_steady=11, **kw)
    with pytest.raises(AssertionError, match="control_freq"):
        pcg.make("PDEControlGym-TrafficPDE1D", control_freq=0, **kw)


def test_fused_autoreset_final_obs(cuda_device):
    """autoreset=True, final_obs=True: info['final_obs'] holds the terminal observation (the oracle's row of the step that
    ended the episode) of the envs that restarted; the returned observation is the initial condition."""
    import torch
    import pdecontrolgym_b200 as pcg
    B, kw = 5, dict(T=2, dt=0.25, X=500, dx=10)
    env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=B, device=cuda_device, simulation_type="outlet", autoreset=True, final_obs=True,
                   reward_class=pcg.TrafficARZReward(), **kw)
    orc = O.TrafficARZOracle(B, simulation_type="outlet", **kw)
    o0, _ = env.reset()
    ic = orc.reset()
    rng = np.random.default_rng(0)
    seen = 0
    for k in range(40):
        a = rng.uniform(1.0, 1.4, size=(B, 1))
        o_r, r_r, te_r, tr_r = orc.step(a)
        o_g, r_g, te_g, tr_g, info = env.step(torch.from_numpy(a))
        done = te_r | tr_r
        assert np.array_equal(te_g.cpu().numpy(), te_r) and np.array_equal(tr_g.cpu().numpy(), tr_r)
        if done.any():
            assert done.all()            # the float time ends every env's episode on the same step (T/dt = 8 -> step 32)
            seen += 1
            assert np.array_equal(info["final_obs"].cpu().numpy(), o_r), f"step {k}: terminal observation"
            assert np.array_equal(o_g.cpu().numpy(), ic), f"step {k}: restarted observation"
            orc.reset()
        else:
            assert np.array_equal(o_g.cpu().numpy(), o_r)
    assert seen >= 1


@pytest.mark.parametrize("st", ["outlet", "both"])
def test_fp32_fast_mode_tolerance_and_event_agreement(cuda_device, st):
    """precision='fp32' (opt-in): float32 state / observation / arithmetic.  Stated tolerance against the float64 mode over
    the whole integrating part of an episode (959 steps, tutorial defaults, actions around qs): observations within 2e-4
    relative L2 per step, rewards within 2e-4 absolute; discrete events (terminated / truncated) must agree on >= 99.9 % of
    env-steps (event agreement rather than trajectory L2 where a threshold can flip)."""
    import torch
    import pdecontrolgym_b200 as pcg
    B = 64
    kw = dict(T=240, dt=0.25, X=500, dx=10, simulation_type=st, limit_pde_state_size=True, reward_class=pcg.TrafficARZReward())
    e64 = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=B, device=cuda_device, precision="fp64", **kw)
    e32 = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=B, device=cuda_device, precision="fp32", **kw)
    o64, _ = e64.reset()
    o32, _ = e32.reset()
    assert o32.dtype == torch.float32 and torch.allclose(o32.double(), o64, rtol=1e-6)
    rng = np.random.default_rng(0)
    nact = 2 if st == "both" else 1
    agree, total, worst_obs, worst_rew = 0, 0, 0.0, 0.0
    for k in range(959):
        a = torch.from_numpy(rng.uniform(0.9 * 1.2, 1.1 * 1.2, (B, nact)))
        o64, r64, te64, tr64, _ = e64.step(a)
        o32, r32, te32, tr32, _ = e32.step(a)
        err = torch.linalg.norm(o32.double() - o64, dim=1) / torch.linalg.norm(o64, dim=1)
        worst_obs = max(worst_obs, float(err.max()))
        worst_rew = max(worst_rew, float((r32 - r64).abs().max()))
        agree += int(((te32 == te64) & (tr32 == tr64)).sum())
        total += B
    assert worst_obs < 2e-4 and worst_rew < 2e-4, (worst_obs, worst_rew)
    assert agree / total >= 0.999, agree / total


@pytest.mark.parametrize("dx", [10, 4])         # 51 points (2 per lane) and 126 points (4 per lane)
def test_exceptional_operands_take_the_exact_division(cuda_device, dx):
    """The substep's divisions run on a deferred-check fast path (csrc/traffic_arz.cu: arz_substep<FAST>); a warp whose
    lanes see a zero / signed-zero / tiny / huge / non-finite operand redoes the substep with the IEEE divisions.  Plant such
    values in the state of some envs (others stay ordinary, so both paths run in one launch) and compare with the oracle."""
    import torch
    import pdecontrolgym_b200 as pcg
    B, kw = 12, dict(T=20, dt=0.25 * dx / 10, X=500, dx=dx)
    env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=B, device=cuda_device, simulation_type="both", control_freq=2,
                   reward_class=pcg.TrafficARZReward(), **kw)
    orc = O.TrafficARZOracle(B, simulation_type="both", control_freq=2, **kw)
    env.reset()
    orc.reset()
    M = orc.M
    plant = [(0, "y", 3, 0.0), (1, "y", 7, -0.0), (2, "r", 10, 1e-150), (3, "y", 5, 1e-305), (4, "r", 20, 1e120),
             (5, "y", M - 2, 0.0), (6, "r", 1, 5e-324), (7, "y", 30, np.inf), (8, "r", 12, np.nan), (9, "r", 25, 0.0)]
    ry = env._ry.cpu().numpy().copy()
    for b, which, j, val in plant:
        if which == "r":
            orc.r[b, j] = val
            ry[b, 0, j] = val
        else:
            orc.y[b, j] = val
            ry[b, 1, j] = val
    env._ry.copy_(torch.from_numpy(ry).to(cuda_device))
    rng = np.random.default_rng(5)
    with np.errstate(all="ignore"):
        for k in range(6):
            a = rng.uniform(0.9, 1.5, size=(B, 2))
            o_r, r_r, te_r, tr_r = orc.step(a)
            o_g, r_g, te_g, tr_g, _ = env.step(torch.from_numpy(a))
            g = o_g.cpu().numpy()
            assert np.array_equal(np.isnan(g), np.isnan(o_r)), f"step {k}: NaN pattern"
            fin = ~np.isnan(o_r)
            assert np.array_equal(g[fin].view(np.uint64), o_r[fin].view(np.uint64)), f"step {k}: obs not bit-identical"
            assert np.array_equal(env._ry.cpu().numpy()[:, 0, :][~np.isnan(orc.r)].view(np.uint64), orc.r[~np.isnan(orc.r)].view(np.uint64))
            assert np.array_equal(te_g.cpu().numpy(), te_r) and np.array_equal(tr_g.cpu().numpy(), tr_r), f"step {k}: flags"


@pytest.mark.parametrize("mapping", ["two_envs_per_warp", "one_env_per_warp"])
@pytest.mark.parametrize("B", [7, 2, 1])
def test_staggered_clocks_and_both_lane_mappings(cuda_device, monkeypatch, mapping, B):
    """Grids up to 64 points can run two envs per warp (16 lanes each, csrc/traffic_arz.cu; opt-in with PDECG_ARZ_LPE16,
    measured slower).  Both must match the oracle, also when the two envs of a warp disagree on `time < T`
    (the integration stops at time_index >= T while the episode runs until T/dt, traffic_arz_env.py:109-111,173) and with
    an odd batch (the last warp's second half shadows the last env)."""
    import torch
    import pdecontrolgym_b200 as pcg
    from pdecontrolgym_b200.envs_traffic_arz import F_TIME
    if mapping == "two_envs_per_warp":
        monkeypatch.setenv("PDECG_ARZ_LPE16", "1")
    else:
        monkeypatch.delenv("PDECG_ARZ_LPE16", raising=False)
    kw = dict(T=5, dt=0.25, X=500, dx=10, simulation_type="outlet", control_freq=2)
    env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=B, device=cuda_device, reward_class=pcg.TrafficARZReward(), **kw)
    orc = O.TrafficARZOracle(B, **kw)
    env.reset()
    orc.reset()
    clocks = np.array([0.0, 4.0, 1.5, 4.5, 3.25, 0.25, 4.75])[:B]      # integration stops at 5.0: envs cross it on different steps
    orc.time_index = clocks.copy()
    env._fenv[:, F_TIME] = torch.from_numpy(clocks).to(cuda_device)
    rng = np.random.default_rng(9)
    mixed = False
    for k in range(30):
        a = rng.uniform(0.9, 1.5, size=(B, 1))
        o_r, r_r, te_r, tr_r = orc.step(a)
        o_g, r_g, te_g, tr_g, _ = env.step(torch.from_numpy(a))
        assert np.array_equal(o_g.cpu().numpy(), o_r), f"step {k}: obs not bit-identical"
        np.testing.assert_allclose(r_g.cpu().numpy(), r_r, rtol=1e-12)
        assert np.array_equal(te_g.cpu().numpy(), te_r) and np.array_equal(tr_g.cpu().numpy(), tr_r)
        assert np.array_equal(env.time_index.cpu().numpy(), orc.time_index)
        mixed |= bool((orc.time_index >= 5.0).any() and (orc.time_index < 5.0).any())
    assert mixed or B == 1


@pytest.mark.parametrize("seed", range(6))
def test_random_physical_parameters(cuda_device, seed):
    """The deferred-check division path must stay bit-identical for any physical parameter set, not only the tutorial's:
    random (vm, rm, tau, steady state, grid, control_freq, simulation type) per case, random in-range actions, 150 steps."""
    import torch
    import pdecontrolgym_b200 as pcg
    rng = np.random.default_rng(100 + seed)
    st = ["inlet", "outlet", "both"][seed % 3]
    dx = float(rng.choice([5.0, 8.0, 10.0, 12.5, 20.0]))
    X = float(rng.choice([300.0, 500.0, 640.0]))
    vm, rm = float(rng.uniform(25, 60)), float(rng.uniform(0.1, 0.3))
    kw = dict(T=30, dt=0.2 * dx / vm * float(rng.uniform(0.6, 1.0)) * 4, X=X, dx=dx, simulation_type=st, control_freq=int(rng.integers(1, 4)),
              v_max=vm, ro_max=rm, tau=float(rng.uniform(20, 120)), ro_steady=float(rng.uniform(0.3, 0.8)) * rm,
              limit_pde_state_size=bool(seed & 1))
    kw["v_steady"] = vm * (1 - kw["ro_steady"] / rm)
    B = 5
    env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=B, device=cuda_device, reward_class=pcg.TrafficARZReward(), **kw)
    orc = O.TrafficARZOracle(B, **kw)
    assert np.array_equal(env.reset()[0].cpu().numpy(), orc.reset())
    qs = kw["ro_steady"] * kw["v_steady"]
    nact = 2 if st == "both" else 1
    with np.errstate(all="ignore"):
        for k in range(150):
            a = rng.uniform(0.75 * qs, 1.25 * qs, size=(B, nact))
            o_r, r_r, te_r, tr_r = orc.step(a)
            o_g, r_g, te_g, tr_g, _ = env.step(torch.from_numpy(a))
            g = o_g.cpu().numpy()
            fin = np.isfinite(o_r)
            assert np.array_equal(np.isnan(g), np.isnan(o_r)) and np.array_equal(g[fin].view(np.uint64), o_r[fin].view(np.uint64)), f"step {k}: obs"
            ok = np.isfinite(r_r)
            np.testing.assert_allclose(r_g.cpu().numpy()[ok], r_r[ok], rtol=1e-12)
            assert np.array_equal(te_g.cpu().numpy(), te_r) and np.array_equal(tr_g.cpu().numpy(), tr_r), f"step {k}: flags"


def test_outlet_train_autoreset_redraws_only_finished_envs(cuda_device):
    """'outlet-train' with autoreset=True: the steady state of an env that finishes is re-drawn on the host with Python's
    `random` (one draw per finishing env, in env order, like the reference's reset(), traffic_arz_env.py:251-256); envs
    that go on are untouched, and steps on which nobody finishes do not reset anything.  Compared with a manual
    step + masked reset driven by the same `random` stream."""
    import random
    import torch
    import pdecontrolgym_b200 as pcg
    from pdecontrolgym_b200.envs_traffic_arz import F_TIME, _RS_VALUES
    B, kw = 6, dict(T=2, dt=0.25, X=500, dx=10, simulation_type="outlet-train", control_freq=2, reward_class=pcg.TrafficARZReward(),
                    rs_init=np.array([0.12, 0.115, 0.125, 0.12, 0.115, 0.125]))
    auto = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=B, device=cuda_device, autoreset=True, **kw)
    manual = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=B, device=cuda_device, autoreset=False, **kw)
    rs0 = np.array([0.125, 0.12, 0.115, 0.115, 0.12, 0.125])
    clocks = torch.tensor([0.0, 6.0, 3.0, 7.5, 1.25, 6.0], dtype=torch.float64, device=cuda_device)   # episodes end at time >= T/dt = 8
    for e in (auto, manual):
        e.reset(options={"rs": rs0})
        e._fenv[:, F_TIME] = clocks
    rng = np.random.default_rng(2)
    acts = rng.uniform(0.9, 1.5, size=(40, B, 1))
    random.seed(11)
    outs = [tuple(t.clone() for t in auto.step(torch.from_numpy(acts[k]))[:4]) for k in range(40)]
    random.seed(11)
    resets = 0
    for k in range(40):
        o, r, te, tr, _ = manual.step(torch.from_numpy(acts[k]))
        done = (te | tr).cpu().numpy()
        if done.any():
            resets += 1
            rs = np.full(B, 0.12)
            rs[np.flatnonzero(done)] = [_RS_VALUES[random.randint(0, 2)] for _ in range(int(done.sum()))]
            o, _ = manual.reset(options={"mask": done, "rs": rs})
        assert torch.equal(outs[k][0], o), f"step {k}: observation"
        assert torch.equal(outs[k][1], r) and torch.equal(outs[k][2], te) and torch.equal(outs[k][3], tr), f"step {k}"
    assert 3 <= resets < 40 and torch.equal(auto._fenv, manual._fenv)

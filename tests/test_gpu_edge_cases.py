"""GPU edge cases shared by all env families: empty batches, same-step auto-reset, masked resets."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
BC = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"], "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}


def test_empty_batches(cuda_device):
    import pdecontrolgym_b200 as pcg
    arz = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=0, device=cuda_device, T=5, dt=0.25, X=500, dx=10)
    o, _ = arz.reset()
    o, r, te, tr, _ = arz.step(np.zeros((0, 1)))
    assert o.shape == (0, 102) and r.shape == (0,)
    bt = pcg.make("PDEControlGym-BrainTumor1D", num_envs=0, device=cuda_device, total_dosage=61.2, T=600, dt=1, X=200, dx=1)
    bt.reset(options={"init_condition": np.zeros(201)})
    o, r, te, tr, info = bt.step(np.zeros(0))
    assert o.shape == (0, 201)
    w = pcg.TherapyWrapper(bt)
    assert w.benchmark().shape == (0,)
    ns = pcg.make("PDEControlGym-NavierStokes2D", num_envs=0, device=cuda_device, boundary_condition=BC, U_ref=np.zeros((200, 21, 21, 2)),
                  action_ref=np.ones(1000), reward_class=pcg.NSReward(), T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05)
    z = np.zeros((21, 21))
    ns.reset(options={"init": (z, z, z)})
    o, r, te, tr, _ = ns.step(np.zeros(0))
    assert o.shape == (0, 21, 21, 2)


def test_arz_autoreset_restarts_finished_envs(cuda_device):
    import torch
    import pdecontrolgym_b200 as pcg
    kw = dict(T=2, dt=0.25, X=500, dx=10, simulation_type="inlet")
    env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=3, device=cuda_device, autoreset=True, **kw)
    ref = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=3, device=cuda_device, **kw)
    o0, _ = env.reset()
    ref.reset()
    a = torch.full((3, 1), 1.3, dtype=torch.float64)
    for k in range(40):
        o, r, te, tr, _ = env.step(a)
        o2, r2, te2, tr2, _ = ref.step(a)
        assert torch.equal(r, r2) and torch.equal(te, te2)
        if te.any():
            assert torch.equal(o[te], o0[te])            # restarted from the initial condition in the same step
            ref.reset(options={"mask": te2})
        else:
            assert torch.equal(o, o2)
    assert k == 39


def test_brain_tumor_autoreset_and_masked_reset(cuda_device):
    import torch
    import pdecontrolgym_b200 as pcg
    xs = np.linspace(0, 60, 61)
    ic = 0.8e5 * np.exp(-0.25 * xs ** 2)
    env = pcg.make("PDEControlGym-BrainTumor1D", num_envs=4, device=cuda_device, total_dosage=61.2, T=30, dt=1, X=60, dx=1, autoreset=True)
    o0, _ = env.reset(options={"init_condition": ic})
    for day in range(30):
        o, r, te, tr, info = env.step(np.zeros(4))
    assert te.all() and torch.equal(o, o0) and (env.time_index == 0).all() and (env.stage == 0).all()
    env.step(np.zeros(4))
    env.reset(options={"init_condition": ic, "mask": np.array([True, False, False, True])})
    assert env.time_index.cpu().tolist() == [0, 1, 1, 0]


def test_navier_stokes_autoreset(cuda_device):
    import torch
    import pdecontrolgym_b200 as pcg
    ns = pcg.make("PDEControlGym-NavierStokes2D", num_envs=2, device=cuda_device, boundary_condition=BC, U_ref=np.zeros((50, 21, 21, 2)),
                  action_ref=np.ones(100), reward_class=pcg.NSReward(), maximum_pressure_iteration=5, T=0.005, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05,
                  autoreset=True)
    z = np.ones((21, 21))
    o0, _ = ns.reset(options={"init": (z, 2 * z, 0 * z)})
    for k in range(4):
        o, r, te, tr, _ = ns.step(np.full(2, 3.0))
    assert te.all() and torch.equal(o, o0) and (ns.time_index == 0).all()
    o, r, te, tr, info = ns.step(np.full(2, 3.0))
    assert info["valid"].all() and not te.any()


def test_state_dict_roundtrip_all_envs(cuda_device):
    """Checkpoint / restore: stepping after load_state_dict() reproduces the original continuation bit for bit."""
    import torch
    import pdecontrolgym_b200 as pcg
    rng = np.random.default_rng(0)
    xs = np.linspace(0, 200, 201)
    z = rng.uniform(-1, 1, (3, 21, 21))
    cases = [
        (lambda: pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=3, device=cuda_device, T=0.02, dt=1e-5, X=1, dx=5e-3,
                          control_sample_rate=1e-3, normalize=True, reward_class=pcg.TunedReward1D(2000)),
         lambda e: e.reset(options={"init_condition": rng.uniform(1, 2, (3, 201)), "beta": np.linspace(0, 9, 201)}),
         lambda k: torch.full((3,), 0.1 * k, device=cuda_device)),
        (lambda: pcg.make("PDEControlGym-TrafficPDE1D", num_envs=3, device=cuda_device, simulation_type="both", T=20, dt=0.25, X=500, dx=10),
         lambda e: e.reset(), lambda k: torch.full((3, 2), 1.2 + 0.01 * k, dtype=torch.float64, device=cuda_device)),
        (lambda: pcg.make("PDEControlGym-BrainTumor1D", num_envs=3, device=cuda_device, total_dosage=61.2, T=600, dt=1, X=200, dx=1),
         lambda e: e.reset(options={"init_condition": 0.8e5 * np.exp(-0.25 * xs ** 2)}), lambda k: torch.full((3,), 0.02, dtype=torch.float64, device=cuda_device)),
        (lambda: pcg.make("PDEControlGym-NavierStokes2D", num_envs=3, device=cuda_device, boundary_condition=BC, U_ref=np.zeros((200, 21, 21, 2)),
                          action_ref=np.ones(1000), reward_class=pcg.NSReward(), maximum_pressure_iteration=30, T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05),
         lambda e: e.reset(options={"init": (z, z * 0.5, z * 0.1)}), lambda k: torch.full((3,), 3.0 - 0.1 * k, dtype=torch.float64, device=cuda_device)),
    ]
    for make, reset, act in cases:
        env = make()
        reset(env)
        for k in range(3):
            env.step(act(k))
        sd = env.state_dict()
        want = [env.step(act(3 + k))[:4] for k in range(3)]
        for k in range(2):
            env.step(act(9))                       # wander off
        env.load_state_dict(sd)
        got = [env.step(act(3 + k))[:4] for k in range(3)]
        for w, g in zip(want, got):
            assert all(torch.equal(a, b) for a, b in zip(w, g))


def test_bc1d_masked_reset_and_recorder(cuda_device):
    import torch
    import pdecontrolgym_b200 as pcg
    from pdecontrolgym_b200.adapters import TrajectoryRecorder
    rng = np.random.default_rng(4)
    N = 6
    kw = dict(T=0.01, dt=1e-5, X=1, dx=5e-3, control_sample_rate=1e-3, normalize=True, reward_class=pcg.NormReward(1000))
    env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=N, device=cuda_device, **kw)
    ic, beta = rng.uniform(1, 2, (N, 201)), rng.uniform(-5, 5, (N, 201))
    rec = TrajectoryRecorder(env, 8)
    rec.reset(options={"init_condition": ic, "beta": beta})
    a = torch.zeros(N, device=cuda_device)
    for k in range(3):
        rec.step(a)
    mask = np.array([True, False, True, False, False, True])
    ic2, beta2 = rng.uniform(1, 2, (N, 201)), rng.uniform(-5, 5, (N, 201))
    obs, _ = env.reset(options={"init_condition": ic2, "beta": beta2, "mask": mask})
    t = env.time_index.cpu().numpy()
    assert t[mask].tolist() == [0, 0, 0] and (t[~mask] == 300).all()
    assert np.array_equal(obs.cpu().numpy()[mask], ic2[mask].astype(np.float32))
    assert np.array_equal(env.beta.cpu().numpy()[mask], beta2[mask]) and np.array_equal(env.beta.cpu().numpy()[~mask], beta[~mask])
    h = rec.history()
    assert h.shape == (4, N, 201) and np.array_equal(h[0].cpu().numpy(), ic.astype(np.float32))
    # the re-seeded envs follow the trajectory of a fresh env with the same IC / beta
    fresh = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=N, device=cuda_device, **kw)
    fresh.reset(options={"init_condition": ic2, "beta": beta2})
    o1, *_ = env.step(a)
    o2, *_ = fresh.step(a)
    assert torch.equal(o1[torch.from_numpy(mask)], o2[torch.from_numpy(mask)])


@pytest.mark.parametrize("env_id,precision", [("PDEControlGym-ReactionDiffusionPDE1D", "fp64"), ("PDEControlGym-TransportPDE1D", "ref")])
@pytest.mark.parametrize("random_ic", [False, True])
@pytest.mark.parametrize("sensing_loc", ["full", "collocated"])
def test_bc1d_fused_autoreset_equals_step_then_masked_reset(cuda_device, env_id, precision, random_ic, sensing_loc):
    """pdecg_bc1d_step_autoreset == pdecg_bc1d_step followed by pdecg_bc1d_reset[_random] on the finished envs."""
    import torch
    import pdecontrolgym_b200 as pcg
    rd = "Reaction" in env_id
    N, steps = 7, 4
    if rd:
        kw = dict(T=steps * 100 * 1e-5, dt=1e-5, X=1, dx=5e-3, control_sample_rate=1e-3)
        npts = 201
    else:
        kw = dict(T=steps * 300 * 1e-4, dt=1e-4, X=1, dx=1e-2, control_sample_rate=3e-2)
        npts = 100
    kw.update(normalize=True, sensing_loc=sensing_loc, precision=precision, limit_pde_state_size=True, max_state_value=1e10,
              reward_class=pcg.TunedReward1D(int(round(kw["T"] / kw["dt"])), -1e3, 3e2))
    rng = np.random.default_rng(9)
    beta = rng.uniform(-3, 3, (N, npts)).astype(np.float64 if precision == "fp64" else np.float32)
    ic = pcg.UniformScaleIC(1.0, 10.0) if random_ic else rng.uniform(1, 2, (N, npts))
    fused = pcg.make(env_id, num_envs=N, device=cuda_device, autoreset=True, **kw)
    manual = pcg.make(env_id, num_envs=N, device=cuda_device, autoreset=False, **kw)
    o1, _ = fused.reset(seed=5, options={"init_condition": ic, "beta": beta})
    o2, _ = manual.reset(seed=5, options={"init_condition": ic, "beta": beta})
    assert torch.equal(o1, o2)
    # stagger the episodes so that only some envs finish at a time
    stag = np.array([True, False, False, True, True, False, True])
    a0 = torch.zeros(N, device=cuda_device)
    fused.step(a0), manual.step(a0)
    fused.reset(options={"mask": stag}), manual.reset(options={"mask": stag})
    n_done = 0
    for k in range(2 * steps + 2):
        a = torch.from_numpy(rng.uniform(-1, 1, N).astype(np.float32))
        o1, r1, te1, tr1, _ = fused.step(a)
        o2, r2, te2, tr2, _ = manual.step(a)
        done = te2 | tr2
        if done.any():
            o2, _ = manual.reset(options={"mask": done})
            n_done += int(done.sum())
        assert torch.equal(te1, te2) and torch.equal(tr1, tr2) and torch.equal(r1, r2), f"step {k}: outputs"
        assert torch.equal(o1, o2), f"step {k}: observation"
        for name in ("u", "time_index", "_carry", "_episode"):
            x, y = getattr(fused, name), getattr(manual, name)
            assert torch.equal(x, y), f"step {k}: {name} differs at {(x != y).nonzero()[:6].tolist()} done={done.tolist()} fused={x[x != y][:6].tolist()} manual={y[x != y][:6].tolist()}"
    assert n_done >= N


def _staggered_pair(fused, manual, reset, act, steps, state_names, stagger):
    """Step a fused-autoreset env and a manual (step, then masked reset) twin through several staggered episodes and
    compare every output and every piece of carried state bit for bit."""
    import torch
    o1, o2 = reset(fused, None), reset(manual, None)
    assert torch.equal(o1, o2)
    stagger(fused), stagger(manual)                     # only some envs finish at a time
    n_done = 0
    for k in range(steps):
        o1, r1, te1, tr1, i1 = fused.step(act(k))
        o2, r2, te2, tr2, i2 = manual.step(act(k))
        done = te2 | tr2
        if done.any():
            if "final_obs" in i1:       # the terminal observation of the restarted envs survives the fused restart
                assert torch.equal(i1["final_obs"][done], o2[done]), f"step {k}: final_obs"
            o2r = reset(manual, done)
            o2 = torch.where(done.reshape((-1,) + (1,) * (o2.dim() - 1)), o2r, o2)
            n_done += int(done.sum())
        assert torch.equal(te1, te2) and torch.equal(tr1, tr2) and torch.equal(r1, r2), f"step {k}: outputs"
        if "valid" in i2:
            assert torch.equal(i1["valid"], i2["valid"])
        assert torch.equal(o1, o2), f"step {k}: observation"
        for name in state_names:
            assert torch.equal(getattr(fused, name), getattr(manual, name)), f"step {k}: {name}"
    return n_done


def _stagger_by_masked_reset(reset, act, mask):
    def go(env):
        for k in range(2):
            env.step(act(k))
        reset(env, mask)
    return go


@pytest.mark.parametrize("simulation_type,dx,vm", [("outlet", 10, 40), ("both", 10, 40), ("inlet", 4, 10)])
def test_arz_fused_autoreset_equals_step_then_masked_reset(cuda_device, simulation_type, dx, vm):
    """pdecg_arz_step_autoreset == pdecg_arz_step followed by pdecg_arz_reset on the finished envs (dx = 4 -> 126
    points, K = 4: the staged row path)."""
    import torch
    import pdecontrolgym_b200 as pcg
    N = 9
    kw = dict(T=1.5, dt=0.25, X=500, dx=dx, simulation_type=simulation_type, v_max=vm, v_steady=vm * (1 - 0.12 / 0.16))   # 24 steps per episode
    fused = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=cuda_device, autoreset=True, final_obs=True, **kw)
    manual = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=cuda_device, autoreset=False, **kw)
    rng = np.random.default_rng(3)
    nact = 2 if simulation_type == "both" else 1
    qs = vm * (1 - 0.12 / 0.16) * 0.12
    acts = [torch.from_numpy(rng.uniform(0.7 * qs, 1.3 * qs, (N, nact))) for _ in range(64)]

    def reset(env, mask):
        return env.reset(options=None if mask is None else {"mask": mask})[0]

    def stagger(env):        # reset() leaves the float time alone (reference quirk), so offset it directly
        env._fenv[:, 0] = torch.arange(N, dtype=torch.float64, device=env.device) % 4 * (3 * 0.25)

    n_done = _staggered_pair(fused, manual, reset, lambda k: acts[k], 60, ("_ry", "_fenv"), stagger)
    assert n_done >= N


def test_brain_tumor_fused_autoreset_equals_step_then_masked_reset(cuda_device):
    """pdecg_bt_step_autoreset == pdecg_bt_run(RAW_STEP) followed by pdecg_bt_reset on the finished envs; per-env
    parameters make some patients die (truncate) and others reach the last day (terminate)."""
    import torch
    import pdecontrolgym_b200 as pcg
    N = 8
    xs = np.linspace(0, 200, 201)
    rng = np.random.default_rng(5)
    ic = 0.8e5 * np.exp(-0.25 * xs ** 2)[None, :] * rng.uniform(0.5, 1.5, (N, 1))
    kw = dict(total_dosage=61.2, T=40, dt=1, X=200, dx=1, D=rng.uniform(0.1, 0.3, N), rho=np.linspace(0.02, 0.5, N),
              t1_detection_radius=3, t1_death_radius=6)
    fused = pcg.make("PDEControlGym-BrainTumor1D", num_envs=N, device=cuda_device, autoreset=True, final_obs=True, **kw)
    manual = pcg.make("PDEControlGym-BrainTumor1D", num_envs=N, device=cuda_device, autoreset=False, **kw)
    acts = [torch.from_numpy(rng.uniform(0, 0.1, N)) for _ in range(130)]

    def reset(env, mask):
        opts = {"init_condition": ic}
        if mask is not None:
            opts["mask"] = mask
        return env.reset(options=opts)[0]

    stag = np.array([True, False, True, False, False, True, False, True])
    n_done = _staggered_pair(fused, manual, reset, lambda k: acts[k], 125, ("u", "_ienv", "_fenv"),
                             _stagger_by_masked_reset(reset, lambda k: acts[k], stag))
    assert n_done >= N


@pytest.mark.parametrize("dx", [0.05, 0.1])
def test_navier_stokes_fused_autoreset_equals_step_then_masked_reset(cuda_device, dx):
    """pdecg_ns_step_autoreset == pdecg_ns_step followed by pdecg_ns_reset on the finished envs, for the warp kernel
    (21 x 21) and the generic shared-memory kernel (11 x 11)."""
    import torch
    import pdecontrolgym_b200 as pcg
    N = 5
    n = int(round(1 / dx + 1))
    rng = np.random.default_rng(7)
    kw = dict(boundary_condition=BC, U_ref=rng.uniform(-1, 1, (50, n, n, 2)), action_ref=np.ones(100), reward_class=pcg.NSReward(0.1),
              maximum_pressure_iteration=7, T=0.006, dt=1e-3, X=1, dx=dx, Y=1, dy=dx)
    fused = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=cuda_device, autoreset=True, final_obs=True, **kw)
    manual = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=cuda_device, autoreset=False, **kw)
    init = tuple(rng.uniform(-1, 1, (N, n, n)) for _ in range(3))
    acts = [torch.from_numpy(rng.uniform(2, 4, N)) for _ in range(20)]

    def reset(env, mask):
        opts = {"init": init}
        if mask is not None:
            opts["mask"] = mask
        return env.reset(options=opts)[0]

    stag = np.array([True, False, True, False, False])
    n_done = _staggered_pair(fused, manual, reset, lambda k: acts[k], 16, ("u", "v", "p", "time_index"),
                             _stagger_by_masked_reset(reset, lambda k: acts[k], stag))
    assert n_done >= N


@pytest.mark.parametrize("autoreset", [False, True])
def test_step_host_matches_step_for_arz_bt_ns(cuda_device, autoreset):
    """pdecg_{arz,bt,ns}_step_host (H2D actions -> kernel -> D2H observation / reward / flags) == the device-resident
    step of a twin env, with and without the fused same-step auto-reset."""
    import torch
    import pdecontrolgym_b200 as pcg
    rng = np.random.default_rng(12)

    def pinned(*shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype).pin_memory()

    # Traffic-ARZ ('both': two actions per env), 24-step episodes
    N = 6
    kw = dict(T=1.5, dt=0.25, X=500, dx=10, simulation_type="both", autoreset=autoreset)
    e1 = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=cuda_device, **kw)
    e2 = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=cuda_device, **kw)
    e1.reset(), e2.reset()
    oh, rh, teh, trh = pinned(N, 102), pinned(N), pinned(N, dtype=torch.uint8), pinned(N, dtype=torch.uint8)
    for k in range(30):
        a = torch.from_numpy(rng.uniform(0.9, 1.5, (N, 2)))
        o, r, te, tr, _ = e1.step(a)
        e2.step_host(a.pin_memory(), oh, rh, teh, trh)
        assert torch.equal(o.cpu(), oh) and torch.equal(r.cpu(), rh), f"arz step {k}"
        assert torch.equal(te.cpu(), teh.bool()) and torch.equal(tr.cpu(), trh.bool())
        assert torch.equal(e1._ry, e2._ry) and torch.equal(e1._fenv, e2._fenv)

    # BrainTumor raw step, 12-day episodes
    xs = np.linspace(0, 100, 101)
    ic = 0.8e5 * np.exp(-0.25 * xs ** 2)[None, :] * rng.uniform(0.5, 1.5, (N, 1))
    kw = dict(total_dosage=61.2, T=12, dt=1, X=100, dx=1, t1_detection_radius=2, autoreset=autoreset)
    e1 = pcg.make("PDEControlGym-BrainTumor1D", num_envs=N, device=cuda_device, **kw)
    e2 = pcg.make("PDEControlGym-BrainTumor1D", num_envs=N, device=cuda_device, **kw)
    e1.reset(options={"init_condition": ic}), e2.reset(options={"init_condition": ic})
    oh, vh = pinned(N, 101), pinned(N, dtype=torch.uint8)
    for k in range(16):
        a = torch.from_numpy(rng.uniform(0, 0.1, N))
        o, r, te, tr, info = e1.step(a)
        e2.step_host(a.pin_memory(), oh, rh, teh, trh, vh)
        assert torch.equal(o.cpu(), oh) and torch.equal(r.cpu(), rh), f"bt day {k}"
        assert torch.equal(te.cpu(), teh.bool()) and torch.equal(tr.cpu(), trh.bool()) and torch.equal(info["valid"].cpu(), vh.bool())
        assert torch.equal(e1._ienv, e2._ienv) and torch.equal(e1._fenv, e2._fenv)

    # NavierStokes (warp kernel), 5-step episodes
    n = 21
    kw = dict(boundary_condition=BC, U_ref=rng.uniform(-1, 1, (50, n, n, 2)), action_ref=np.ones(100), reward_class=pcg.NSReward(0.1),
              maximum_pressure_iteration=9, T=0.006, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05, autoreset=autoreset)
    e1 = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=cuda_device, **kw)
    e2 = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=cuda_device, **kw)
    init = tuple(rng.uniform(-1, 1, (N, n, n)) for _ in range(3))
    e1.reset(options={"init": init}), e2.reset(options={"init": init})
    oh = pinned(N, n, n, 2)
    for k in range(8):
        a = torch.from_numpy(rng.uniform(2, 4, N))
        o, r, te, tr, info = e1.step(a)
        e2.step_host(a.pin_memory(), oh, rh, teh, trh, vh)
        assert torch.equal(o.cpu(), oh) and torch.equal(r.cpu(), rh), f"ns step {k}"
        assert torch.equal(te.cpu(), teh.bool()) and torch.equal(info["valid"].cpu(), vh.bool())
        assert torch.equal(e1.p, e2.p) and torch.equal(e1.time_index, e2.time_index)


def test_therapy_wrapper_masked_reset_keeps_the_other_patients(cuda_device):
    """TherapyWrapper.reset(options={'mask': ...}) restarts only the selected patients: the others -- mid-therapy, with
    weekends=True -- keep their consecutive-treatment-day counter, so their forced rest days land where the oracle's do."""
    import torch
    import pdecontrolgym_b200 as pcg
    from oracle import np_oracle as O
    from tests.golden.gen_golden import BT_P, bt_ic
    N = 4
    ic = bt_ic(200, 201)
    env = pcg.make("PDEControlGym-BrainTumor1D", num_envs=N, device=cuda_device, reward_class=pcg.BrainTumorReward(),
                   reset_init_condition_func=bt_ic, **BT_P)
    w = pcg.TherapyWrapper(env, weekends=True)
    oenvs = [O.BrainTumorOracle(1, **BT_P) for _ in range(N)]
    ows = [O.TherapyWrapperOracle(e, ic, weekends=True) for e in oenvs]
    tb = w.benchmark()
    for ow in ows:
        assert ow.benchmark() == int(tb[0])
    o, _ = w.reset(options={"init_condition": ic})
    refs = [ow.reset()[0] for ow in ows]
    assert all(np.array_equal(o[i].cpu().numpy(), refs[i]) for i in range(N))
    mask = np.array([True, False, True, False])
    for k in range(9):
        if k == 3:                        # three treatment days in: patients 0 and 2 start over, 1 and 3 carry on
            o, _ = w.reset(options={"init_condition": ic, "mask": mask})
            for i in np.nonzero(mask)[0]:
                assert np.array_equal(o[i].cpu().numpy(), ows[i].reset()[0])
            assert w.consecutive_treatment_days.cpu().tolist() == [0, 3, 0, 3]
        o, r, te, tr, _ = w.step(torch.full((N,), 0.02, dtype=torch.float64))
        for i in range(N):
            o2, r2, te2, tr2 = ows[i].step(0.02)
            np.testing.assert_allclose(o[i].cpu().numpy(), o2[0], rtol=1e-13, atol=0, err_msg=f"call {k} env {i}")
            assert float(r[i]) == float(r2[0]) and bool(te[i]) == bool(te2[0]) and bool(tr[i]) == bool(tr2[0])
        assert env.time_index.cpu().tolist() == [int(e.t[0]) for e in oenvs], f"call {k}: day counters (weekend placement)"


def test_nonfinite_flag_and_optional_truncation(cuda_device):
    """ReactionDiffusion with Neumann control and normalize=True overflows to inf within a few steps (reference quirk Q4,
    parabolic.py:148-150).  info['nonfinite'] flags exactly the envs whose state row holds an inf / NaN (the oracle's
    np.isfinite), and truncate_nonfinite=True ends their episodes (an opt-in extension: the reference lets inf / NaN run on)."""
    import torch
    import pdecontrolgym_b200 as pcg
    from oracle import np_oracle as O
    from tests.helpers import bc1d_case, make_ics
    B, steps = 6, 6
    kw, npts, beta, nt = bc1d_case("rd", None, steps)
    common = dict(normalize=True, control_type="Neumann", max_control_value=20, **kw)
    rng = np.random.default_rng(8)
    ics = make_ics(rng, B, npts)
    ics[::2] *= 1e-30                                   # every other env stays tiny for a while longer
    orc = O.ReactionDiffusionOracle(B, reward=O.NormRewardSpec(nt - 1, norm=2), precision="ref", **common)
    env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=B, device=cuda_device, precision="ref", reward_class=pcg.NormReward(nt - 1, norm="2"),
                   report_nonfinite=True, **common)
    env_t = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=B, device=cuda_device, precision="ref", reward_class=pcg.NormReward(nt - 1, norm="2"),
                     truncate_nonfinite=True, **common)
    orc.reset(ics, beta.astype(np.float32))
    for e in (env, env_t):
        e.reset(options={"init_condition": ics, "beta": beta.astype(np.float32)})
    seen = False
    for k in range(steps - 1):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        o_r, r_r, te_r, tr_r = orc.step(a)
        o, r, te, tr, info = env.step(torch.from_numpy(a))
        o2, r2, te2, tr2, info2 = env_t.step(torch.from_numpy(a))
        bad = ~np.isfinite(o_r).all(axis=1)
        assert np.array_equal(info["nonfinite"].cpu().numpy(), bad), f"step {k}"
        assert not tr.any() and np.array_equal(tr2.cpu().numpy(), bad)
        assert np.array_equal(o.cpu().numpy(), o_r, equal_nan=True)
        seen |= bool(bad.any())
    assert seen

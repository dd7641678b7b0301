"""CUDA parity: BrainTumor1D / TherapyWrapper kernels vs the numpy oracle + the reference's published
day counts (K3 / K4 of SURVEY.md §4: BrainTumor1DExample.ipynb:11745,13104,14084,14649,15682,16214).

Every day -- Growth, Therapy, Post-Therapy -- is BIT-IDENTICAL to the oracle evaluated with the correctly rounded exp()
(`exp="cr"`: oracle/c restatement of the kernels' double-double exp_cr, pinned to mpmath in tests/test_exp_cr.py).  Against
the oracle with numpy's own exp (what the reference calls; its last bit depends on the host CPU, see
tests/test_exp_cr.py::test_numpy_exp_is_not_a_function_of_its_argument_alone) therapy-stage states agree to 1e-13 relative
and every discrete event exactly."""
import numpy as np
import pytest

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

P = dict(T=600, X=200, dt=1, dx=1, dosage_termination_threshold=0.1, t1_detection_threshold=0.8,
         t2_detection_threshold=0.16, alpha=0.04, alpha_beta_ratio=10, k=1e5, t1_detection_radius=15,
         t1_death_radius=35, total_dosage=61.2)


def ic_fn(X, nx):
    xs = np.linspace(0, X, nx)
    return 0.8 * 1e5 * np.exp(-0.25 * (xs ** 2))


def make_gpu(cuda_device, B, D=0.2, rho=0.03, **over):
    import pdecontrolgym_b200 as pcg
    kw = {**P, **over}
    return pcg.make("PDEControlGym-BrainTumor1D", num_envs=B, device=cuda_device, D=D, rho=rho,
                    reward_class=pcg.BrainTumorReward(), reset_init_condition_func=ic_fn, **kw)


@pytest.mark.parametrize("exp_mode", ["cr", "numpy"])
def test_raw_env_vs_oracle(cuda_device, exp_mode):
    import torch
    B = 6
    rng = np.random.default_rng(0)
    D = 0.2 * rng.uniform(0.9, 1.1, B)
    rho = 0.03 * rng.uniform(0.9, 1.1, B)
    env = make_gpu(cuda_device, B, D=D, rho=rho)
    orc = O.BrainTumorOracle(B, D=D, rho=rho, exp=exp_mode, **P)
    tb = np.array([363, 350, 370, 363, 340, 380])
    env.t_benchmark = tb
    orc.t_benchmark = tb
    o_g, _ = env.reset()
    o_r = orc.reset(ic_fn(200, 201))
    assert np.array_equal(o_g.cpu().numpy(), o_r)
    exact_days = 0
    for day in range(603):
        a = rng.uniform(0, 0.12, B)
        a[rng.uniform(size=B) < 0.3] = 0.0
        o_r, r_r, te_r, tr_r, valid_r = orc.step(a)
        o_g, r_g, te_g, tr_g, info = env.step(torch.from_numpy(a))
        og = o_g.cpu().numpy()
        live = valid_r
        assert np.array_equal(info["valid"].cpu().numpy(), valid_r)
        assert np.array_equal(info["stage"].cpu().numpy()[live], orc.stage[live]), f"day {day}: stage"
        if exp_mode == "cr":
            assert np.array_equal(og, o_r), f"day {day}: state not bit-identical (max rel {np.max(np.abs(og - o_r) / np.maximum(np.abs(o_r), 1e-300))})"
        else:
            np.testing.assert_allclose(og, o_r, rtol=1e-13, atol=0, err_msg=f"day {day}")
        exact_days += int(np.array_equal(og, o_r))
        assert np.array_equal(te_g.cpu().numpy()[live], te_r[live]) and np.array_equal(tr_g.cpu().numpy()[live], tr_r[live]), f"day {day}: flags"
        np.testing.assert_allclose(r_g.cpu().numpy()[live], r_r[live], rtol=1e-10, atol=1e-12, err_msg=f"day {day}: reward")
        for name in ("growthDays", "therapyDays", "postTherapyDays", "simulationDays", "firstTherapyDay",
                     "firstPostTherapyDay", "cDeathDay"):
            assert np.array_equal(getattr(env, name).cpu().numpy(), getattr(orc, name)), f"day {day}: {name}"
        np.testing.assert_allclose(env.remaining_dosage.cpu().numpy(), orc.remaining, rtol=1e-15)
    assert not valid_r.any()          # every env ran past the end of its episode (t == nt-1)
    # numpy's exp: bit-identical at least until the first env of the batch (jittered D / rho) enters Therapy
    assert exact_days >= (603 if exp_mode == "cr" else int(orc.growthDays.min()) - 1)


@pytest.mark.parametrize("schedule,weekends,first_post,death,total", [
    ("open", False, -1, 363, 0.0),        # K3
    ("paper", False, 248, 430, 67.0),     # K4 (no weekends)
    ("paper", True, 260, 431, 68.0),      # K4 (weekend breaks)
])
def test_wrapper_known_answers(cuda_device, schedule, weekends, first_post, death, total):
    import pdecontrolgym_b200 as pcg
    B = 3
    env = make_gpu(cuda_device, B)
    w = pcg.TherapyWrapper(env, weekends=weekends)
    tb = w.benchmark()
    assert tb.cpu().tolist() == [363] * B
    orc = O.BrainTumorOracle(1, exp="cr", **P)
    ow = O.TherapyWrapperOracle(orc, ic_fn(200, 201), weekends=weekends)
    assert ow.benchmark() == 363
    obs, info = w.reset()
    o_r = ow.reset()
    assert np.array_equal(obs.cpu().numpy()[1], o_r[0])
    assert env.growthDays.cpu().tolist() == [213] * B and env.firstTherapyDay.cpu().tolist() == [214] * B
    props = [1 / 34] * 34 if schedule == "paper" else []
    ret = np.zeros(B)
    k = 0
    while True:
        a = props[k] if k < len(props) else 0.0
        k += 1
        obs, r, te, tr, info = w.step(np.full(B, a))
        o_r, r_r, te_r, tr_r = ow.step(a)
        assert np.array_equal(obs.cpu().numpy()[2], o_r[0]), f"call {k}: therapy-stage state not bit-identical"
        np.testing.assert_allclose(r.cpu().numpy(), np.full(B, r_r[0]), rtol=1e-10, atol=1e-12)
        assert te.cpu().tolist() == [bool(te_r[0])] * B and tr.cpu().tolist() == [bool(tr_r[0])] * B
        ret += r.cpu().numpy()
        if te_r[0] or tr_r[0]:
            break
    assert env.cDeathDay.cpu().tolist() == [death] * B
    assert env.firstPostTherapyDay.cpu().tolist() == [first_post] * B
    assert ret.tolist() == [total] * B
    assert w.soft_constraint_violations.cpu().tolist() == [ow.soft_constraint_violations] * B
    assert w.treatment_calls.cpu().tolist() == [ow.treatment_calls] * B


def test_wrapper_random_schedules_with_violations(cuda_device):
    """Per-env dose schedules incl. over-dosing (negative toxicity rewards) and weekend breaks."""
    import pdecontrolgym_b200 as pcg
    B = 5
    rng = np.random.default_rng(7)
    sched = rng.uniform(0, 0.25, (B, 80))
    sched[rng.uniform(size=sched.shape) < 0.25] = 0.0
    env = make_gpu(cuda_device, B)
    w = pcg.TherapyWrapper(env, weekends=True)
    w.benchmark()
    w.reset()
    done = np.zeros(B, dtype=bool)
    outs = []
    for k in range(80):
        obs, r, te, tr, info = w.step(sched[:, k])
        outs.append((obs.cpu().numpy(), r.cpu().numpy(), te.cpu().numpy(), tr.cpu().numpy()))
    for b in range(B):
        orc = O.BrainTumorOracle(1, exp="cr", **P)
        ow = O.TherapyWrapperOracle(orc, ic_fn(200, 201), weekends=True)
        ow.benchmark()
        ow.reset()
        for k in range(80):
            o_r, r_r, te_r, tr_r = ow.step(sched[b, k])
            assert np.array_equal(outs[k][0][b], o_r[0]), f"env {b} step {k}: state not bit-identical"
            np.testing.assert_allclose(outs[k][1][b], r_r[0], rtol=1e-10, atol=1e-12, err_msg=f"env {b} step {k}")
            assert outs[k][2][b] == te_r[0] and outs[k][3][b] == tr_r[0]
            if te_r[0] or tr_r[0]:
                break
        assert ow.soft_constraint_violations > 0
        assert int(w.soft_constraint_violations[b]) >= ow.soft_constraint_violations


@pytest.mark.parametrize("X", [20, 63, 100, 255, 256, 300, 400, 460, 511])
def test_grid_sizes(cuda_device, X):
    """All points-per-lane instantiations; small domains (tumour reaches the wall)."""
    import torch
    B = 2
    over = dict(T=120, X=X)
    env = make_gpu(cuda_device, B, **over)
    orc = O.BrainTumorOracle(B, **{**P, **over})
    env.t_benchmark = [50, 60]
    orc.t_benchmark = np.array([50, 60])
    env.reset()
    orc.reset(ic_fn(X, X + 1))
    for day in range(120):
        a = np.full(B, 0.03)
        o_r, r_r, te_r, tr_r, v = orc.step(a)
        o_g, r_g, te_g, tr_g, info = env.step(torch.from_numpy(a))
        np.testing.assert_allclose(o_g.cpu().numpy(), o_r, rtol=1e-13, atol=0)
        assert np.array_equal(te_g.cpu().numpy()[v], te_r[v]) and np.array_equal(tr_g.cpu().numpy()[v], tr_r[v])


def test_fp32_fast_mode_tolerance_and_event_agreement(cuda_device):
    """precision='fp32' (opt-in): float32 state / observation / finite-difference arithmetic.  Against the float64 mode over
    whole episodes of 64 patients with jittered D / rho and a shared random dose schedule: densities within 2e-3 of the
    carrying capacity (absolute; 5e-4 typical), and the discrete events -- stage switches, death / last day -- land on the
    same day for >= 95 % of the patients and within 2 days for all of them (a threshold crossing can move by a day when the
    density at the detection radius differs in the 7th digit: SURVEY 8(c) asks for the agreement rate, not a trajectory L2)."""
    import torch
    B = 64
    rng = np.random.default_rng(1)
    D = 0.2 * rng.uniform(0.9, 1.1, B)
    rho = 0.03 * rng.uniform(0.9, 1.1, B)
    e64 = make_gpu(cuda_device, B, D=D, rho=rho, precision="fp64")
    e32 = make_gpu(cuda_device, B, D=D, rho=rho, precision="fp32")
    for e in (e64, e32):
        e.t_benchmark = np.full(B, 363)
        e.reset()
    assert e32.u.dtype == torch.float32
    worst = 0.0
    for day in range(600):
        a = torch.full((B,), 1 / 34 if day % 7 < 5 else 0.0, dtype=torch.float64)
        o64, r64, te64, tr64, i64 = e64.step(a)
        o32, r32, te32, tr32, i32 = e32.step(a)
        live = i64["valid"] & i32["valid"]
        if live.any():
            worst = max(worst, float((o32.double() - o64).abs()[live].max()) / 1e5)
    assert worst < 2e-3, worst
    for name, tol_days in (("firstTherapyDay", 2), ("firstPostTherapyDay", 2), ("cDeathDay", 2), ("simulationDays", 2)):
        a, b = getattr(e64, name).cpu().numpy(), getattr(e32, name).cpu().numpy()
        assert np.abs(a - b).max() <= tol_days, (name, np.abs(a - b).max())
        assert (a == b).mean() >= 0.95, (name, (a == b).mean())

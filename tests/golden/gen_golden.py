"""Regenerates tests/golden/*.npz by running the UNMODIFIED reference (lukebhan/PDEControlGym at
/root/reference, numpy 2.3.5) through oracle/ref_loader.py.  Only runnable in the build container
(the GPU box has no /root/reference); the produced fixtures are committed.

    python tests/golden/gen_golden.py
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.simplefilter("ignore")

from oracle.ref_loader import load_reference, make_shimmed_norm_reward  # noqa: E402
from tests.helpers import bc1d_case, make_ics  # noqa: E402


class _LazyRef:
    """The reference is imported on first use so that test modules can import the case tables below on
    machines without /root/reference (the GPU box)."""

    def __getattr__(self, name):
        return getattr(load_reference(), name)


ref = _LazyRef()
quiet = lambda: contextlib.redirect_stdout(io.StringIO())  # noqa: E731

BC1D_CASES = [
    # name, env, control, sensing_loc, sensing_type, normalize, beta dtype, action dtype, reward, f64 state
    ("rd_dir_full_tuned", "rd", "Dirchilet", "full", None, True, "float32", "float32", "tuned", False),
    ("rd_dir_full_norm_b64", "rd", "Dirchilet", "full", None, True, "float64", "float32", ("temporal", 2), False),
    ("rd_neu_col_norm", "rd", "Neumann", "collocated", None, False, "float64", "float64", ("temporal", 2), False),
    ("rd_dir_opp_diff", "rd", "Dirchilet", "opposite", "Neumann", True, "float32", "float32", ("differential", 1), False),
    ("rd_dir_full_norm_fp64", "rd", "Dirchilet", "full", None, True, "float64", "float32", ("temporal", 2), True),
    ("tr_dir_full_tuned", "tr", "Dirchilet", "full", None, True, "float32", "float32", "tuned", False),
    ("tr_dir_full_tuned_b64", "tr", "Dirchilet", "full", None, False, "float64", "float64", "tuned", False),
    ("tr_neu_opp_norm", "tr", "Neumann", "opposite", "Dirchilet", True, "float32", "float32", ("temporal", np.inf), False),
    ("tr_dir_col_tuned_fp64", "tr", "Dirchilet", "collocated", None, True, "float64", "float32", "tuned", True),
]


# grids beyond the examples' (the CTA-per-env CUDA kernels): name, ..., f64 state, nx, control steps
BC1D_LARGE_CASES = [
    ("rd_700_dir_full_tuned", "rd", "Dirchilet", "full", None, True, "float32", "float32", "tuned", False, 700, 4),
    ("tr_1500_dir_full_tuned", "tr", "Dirchilet", "full", None, True, "float32", "float32", "tuned", False, 1500, 4),
    ("rd_2500_neu_col_norm_fp64", "rd", "Neumann", "collocated", None, False, "float64", "float64", ("temporal", 2), True, 2500, 3),
    ("tr_4097_dir_opp_diff_b64", "tr", "Dirchilet", "opposite", "Neumann", True, "float64", "float32", ("differential", 1), False, 4097, 3),
]


def gen_bc1d(cases=None, fname="bc1d_reference_runs.npz"):
    out = {}
    for case in (BC1D_CASES if cases is None else cases):
        (name, env, ct, sl, st, nz, bd, ad, rk, f64) = case[:10]
        nx, steps = (case[10], case[11]) if len(case) > 10 else (None, 12)
        kw, npts, beta, nt = bc1d_case(env, nx, steps)
        beta = beta.astype(bd)
        nt_r = nt - 1
        if rk == "tuned":
            rr = ref.TunedReward1D(nt_r, -1e3, 3e2)
        else:
            rr = make_shimmed_norm_reward(ref, nt=nt_r, norm=rk[1], horizon=rk[0])
        rng = np.random.default_rng(abs(hash(name)) % 2**32 if False else sum(map(ord, name)))
        ic = make_ics(rng, 1, npts)[0]
        Ref = ref.ReactionDiffusionPDE1D if env == "rd" else ref.TransportPDE1D
        renv = Ref(reward_class=rr, sensing_noise_func=lambda s: s, reset_init_condition_func=lambda nx: ic,
                   reset_recirculation_func=lambda nx: beta, normalize=nz, sensing_loc=sl, control_type=ct,
                   sensing_type=st, limit_pde_state_size=False, max_state_value=30.0, max_control_value=20, **kw)
        obs0, _ = renv.reset()
        if f64:   # the "fp64" precision mode: same code, state buffer re-created as float64
            renv.u = np.zeros(renv.u.shape, dtype=np.float64)
            renv.u[0] = ic
            renv.beta = np.asarray(renv.beta, dtype=np.float64)
            obs0 = renv.sensing_update(renv.u[0], renv.dx, renv.sensing_noise_func)
        acts, obs, rew, term, trunc = [], [np.asarray(obs0)], [], [], []
        for k in range(steps):
            a = np.array([rng.uniform(-1, 1) if nz else rng.uniform(-3, 3)], dtype=ad)
            o, r, te, tr, _ = renv.step(a)
            acts.append(a[0]); obs.append(np.array(o)); rew.append(float(r)); term.append(bool(te)); trunc.append(bool(tr))
        out[name + "/ic"] = ic
        out[name + "/beta"] = beta
        out[name + "/actions"] = np.array(acts, dtype=ad)
        out[name + "/obs"] = np.array(obs)
        out[name + "/reward"] = np.array(rew)
        out[name + "/terminated"] = np.array(term)
        out[name + "/truncated"] = np.array(trunc)
    np.savez_compressed(os.path.join(HERE, fname), **out)
    print(fname, len(BC1D_CASES if cases is None else cases), "cases")


def gen_arz():
    out = {}
    for st in ("inlet", "outlet", "both"):
        for cf in (1, 2):
            with quiet():
                env = ref.TrafficPDE1D(simulation_type=st, control_freq=cf, limit_pde_state_size=True,
                                       reward_class=ref.TrafficARZReward(), T=5, dt=0.25, X=500, dx=10)
            o0, _ = env.reset()
            rng = np.random.default_rng(len(st) * 10 + cf)
            nact = 2 if st == "both" else 1
            acts, obs, rew, term, trunc = [], [o0.copy()], [], [], []
            for k in range(100):
                a = rng.uniform(0.7 * 1.2, 1.3 * 1.2, size=nact)
                o, r, te, tr, _ = env.step(a.copy())
                acts.append(a); obs.append(o.copy()); rew.append(float(r)); term.append(bool(te)); trunc.append(bool(tr))
            key = f"{st}_cf{cf}"
            out[key + "/actions"] = np.array(acts)
            out[key + "/obs"] = np.array(obs)[::5]
            out[key + "/reward"] = np.array(rew)
            out[key + "/terminated"] = np.array(term)
            out[key + "/truncated"] = np.array(trunc)
    np.savez_compressed(os.path.join(HERE, "arz_reference_runs.npz"), **out)
    print("arz: 6 cases")


BT_P = dict(T=600, X=200, dt=1, dx=1, dosage_termination_threshold=0.1, t1_detection_threshold=0.8,
            t2_detection_threshold=0.16, D=0.2, rho=0.03, alpha=0.04, alpha_beta_ratio=10, k=1e5,
            t1_detection_radius=15, t1_death_radius=35, total_dosage=61.2)


def bt_ic(X, nx):
    xs = np.linspace(0, X, nx)
    return 0.8 * 1e5 * np.exp(-0.25 * (xs ** 2))


def gen_brain_tumor():
    out = {}
    # raw env, constant action 0.05, t_benchmark = 363  (pin S5)
    env = ref.BrainTumor1D(reward_class=ref.BrainTumorReward(), normalize=True, reset_init_condition_func=bt_ic,
                           verbose=False, **BT_P)
    with quiet():
        env.reset()
    env.t_benchmark = 363
    obs, rew, term, trunc, stage = [], [], [], [], []
    while True:
        with quiet():
            o, r, te, tr, info = env.step(0.05)
        obs.append(o.copy()); rew.append(float(r)); term.append(bool(te)); trunc.append(bool(tr))
        stage.append(["Growth", "Therapy", "Post-Therapy"].index(info["stage"]))
        if te or tr:
            break
    days = np.array([1, 100, 213, 214, 220, 234, 235, 300, 400, len(obs)]) - 1
    out["raw/obs_days"] = days + 1
    out["raw/obs"] = np.array(obs)[days]
    out["raw/reward"] = np.array(rew)
    out["raw/terminated"] = np.array(term)
    out["raw/truncated"] = np.array(trunc)
    out["raw/stage"] = np.array(stage)
    # wrapper schedules (K3 / K4)
    for name, props, weekends in (("open", [], False), ("paper", [1 / 34] * 34, False), ("paper_weekends", [1 / 34] * 34, True)):
        env = ref.BrainTumor1D(reward_class=ref.BrainTumorReward(), normalize=True, reset_init_condition_func=bt_ic,
                               verbose=False, **BT_P)
        w = ref.TherapyWrapper(env, weekends=weekends, verbose=False)
        with quiet():
            tb = w.benchmark()
            o, _ = w.reset()
        obs, rew, k = [o.copy()], [], 0
        te = tr = False
        while not (te or tr):
            a = props[k] if k < len(props) else 0
            k += 1
            with quiet():
                o, r, te, tr, info = w.step(a)
            obs.append(o.copy()); rew.append(float(r))
        out[f"wrap_{name}/summary"] = np.array([tb, env.growthDays, env.firstTherapyDay, -1 if env.firstPostTherapyDay is None else env.firstPostTherapyDay,
                                                env.cDeathDay, len(rew)])
        out[f"wrap_{name}/reward"] = np.array(rew)
        out[f"wrap_{name}/obs_first_last"] = np.array([obs[0], obs[1], obs[-1]])
    np.savez_compressed(os.path.join(HERE, "brain_tumor_reference_runs.npz"), **out)
    print("brain tumor: raw +", 3, "wrapper schedules")


def gen_ns():
    tg = np.load(os.path.join(ref_root(), "examples", "NavierStokes", "target.npz"))
    rows = np.array([1, 2, 10, 50, 100, 150, 199])
    np.savez_compressed(os.path.join(HERE, "ns_target_subset.npz"), rows=rows, u=tg["u"][rows], v=tg["v"][rows], u0=tg["u"][0])
    # short run with mixed boundary conditions and few sweeps
    Uref = np.stack([tg["u"], tg["v"]], axis=-1)
    bc = {"upper": ["Controllable", "Neumann"], "lower": ["Neumann", "Dirchilet"], "left": ["Dirchilet", "Controllable"], "right": ["Neumann", "Neumann"]}
    rng = np.random.default_rng(5)
    u0, v0, p0 = (rng.uniform(-1, 1, (21, 21)) for _ in range(3))
    env = ref.NavierStokes2D(reset_init_condition_func=lambda X: (u0.copy(), v0.copy(), p0.copy()), boundary_condition=bc,
                             U_ref=Uref, action_ref=2.0 * np.ones(1000), reward_class=ref.NSReward(0.1), action_dim=1,
                             maximum_pressure_iteration=37, T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05)
    env.reset()
    acts, obs, rew, ps = [], [], [], []
    for k in range(6):
        a = rng.uniform(2, 4)
        o, r, te, tr, _ = env.step(a)
        acts.append(a); obs.append(o.copy()); rew.append(float(r)); ps.append(env.p.copy())
    np.savez_compressed(os.path.join(HERE, "ns_reference_run.npz"), u0=u0, v0=v0, p0=p0, actions=np.array(acts), obs=np.array(obs),
                        reward=np.array(rew), p=np.array(ps), uref_rows=Uref[:8])
    print("ns: target subset + mixed-BC run")


def gen_ns_adjoint():
    """examples/NavierStokes/NS2Doptimization.py:62-107 run with the REFERENCE's own env, central_difference, laplace and
    solve_pressure (same statements, same order), shortened to T = 12 recorded steps and 60 Jacobi sweeps so the fixture stays
    small: forward rollout with random actions, backward costate sweep, new action sequence."""
    tg = np.load(os.path.join(ref_root(), "examples", "NavierStokes", "target.npz"))
    u_target, v_target = tg["u"], tg["v"]
    desire_states = np.stack([u_target, v_target], axis=-1)
    bc = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"], "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}
    rng = np.random.default_rng(17)
    c = rng.uniform(-5, 5, 3)
    dt, dx, dy = 1e-3, 0.05, 0.05
    env = ref.NavierStokes2D(reset_init_condition_func=lambda X: tuple(ci * np.ones_like(X) for ci in c), boundary_condition=bc,
                             U_ref=desire_states, action_ref=2.0 * np.ones(1000), reward_class=ref.NSReward(0.1), action_dim=1,
                             maximum_pressure_iteration=60, T=0.2, dt=dt, X=1, dx=dx, Y=1, dy=dy)
    central_difference, laplace = ref.central_difference, ref.laplace

    def apply_boundary(a1, a2):                 # :55-60
        a1[:, [-1, 0]] = 0.
        a1[[-1, 0], :] = 0.
        a2[:, [-1, 0]] = 0.
        a2[[-1, 0], :] = 0.
        return a1, a2
    T = 12
    env.reset(seed=400)
    U, V, acts_in = [], [], []
    for t in range(T):                          # :68-72
        a = rng.uniform(2, 4)
        acts_in.append(a)
        env.step(a)
        U.append(env.u)
        V.append(env.v)
    u_ref = [2 for _ in range(T)]
    Lam1, Lam2 = [], []                          # :79-101
    Lam1.append(np.zeros_like(U[0]))
    Lam2.append(np.zeros_like(U[0]))
    pressure = np.zeros_like(U[0])
    for t in range(T - 1):
        lam1, lam2 = Lam1[-1], Lam2[-1]
        dl1dx, dl1dy = central_difference(lam1, "x", dx), central_difference(lam1, "y", dy)
        dl2dx, dl2dy = central_difference(lam2, "x", dx), central_difference(lam2, "y", dy)
        laplace_l1, laplace_l2 = laplace(lam1, dx, dy), laplace(lam2, dx, dy)
        dlam1dt = - 2 * dl1dx * U[-1-t] - dl1dy * V[-1-t] - dl2dx * V[-1-t] - 0.1 * laplace_l1 + (U[-1-t]-u_target[-1-t])
        dlam2dt = - 2 * dl2dy * V[-1-t] - dl1dy * U[-1-t] - dl2dx * U[-1-t] - 0.1 * laplace_l2 + (V[-1-t]-v_target[-1-t])
        lam1 = lam1 - dt * dlam1dt
        lam2 = lam2 - dt * dlam2dt
        lam1, lam2 = apply_boundary(lam1, lam2)
        pressure = env.solve_pressure(lam1, lam2, pressure)
        lam1 = lam1 - dt * central_difference(pressure, "x", dx)
        lam2 = lam2 - dt * central_difference(pressure, "y", dy)
        lam1, lam2 = apply_boundary(lam1, lam2)
        Lam1.append(lam1)
        Lam2.append(lam2)
    Lam1 = Lam1[::-1]
    actions = []                                 # :103-106
    for t in range(T):
        dl1dx2 = central_difference(Lam1[t], "y", dy)
        actions.append(u_ref[t] - 0.1/0.1 * sum(dl1dx2[-2, :])*5*dx)
    rows = np.arange(len(u_target) - T - 1, len(u_target))          # the target rows the backward sweep indexes from the end
    np.savez_compressed(os.path.join(HERE, "ns_adjoint_reference_run.npz"), init=c, actions_in=np.array(acts_in), U=np.array(U), V=np.array(V),
                        Lam1=np.array(Lam1), Lam2=np.array(Lam2[::-1]), actions_out=np.array(actions), target_rows=rows,
                        u_target_tail=u_target[rows], v_target_tail=v_target[rows], uref_head=desire_states[:T + 2])
    print("ns adjoint: T =", T)


def ref_root():
    from oracle.ref_loader import REFERENCE_ROOT
    return REFERENCE_ROOT


if __name__ == "__main__":
    gen_bc1d()
    gen_bc1d(BC1D_LARGE_CASES, "bc1d_large_grid_reference_runs.npz")
    gen_arz()
    gen_brain_tumor()
    gen_ns()
    gen_ns_adjoint()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))

"""CPU arm of bench.py -- TEST / BASELINE INFRASTRUCTURE ONLY (never imported by pdecontrolgym_b200/).

The reference's own CPU path for every BASELINE.json configuration: the UNMODIFIED env / reward classes of
lukebhan/PDEControlGym (staged by oracle/make_ref.py, imported through oracle/ref_loader.py) stepped the way
gymnasium's AsyncVectorEnv steps them -- one worker process per host core, `k` single-env instances per worker,
lock-stepped through pipes with same-step auto-reset.  When no reference tree can be found the numpy oracle
(oracle/np_oracle.py, the port) takes its place and the result says `kind: "port"`.

Workloads: "transport1" (config 1), "rd" (config 2, the headline), "arz" and "bt" (config 3),
"ns" (config 4).  Units of work per env-step are in `PTS`.
"""
import contextlib
import io
import os
import time
import warnings

import numpy as np

PTS = {"transport1": 99 * 1000, "rd": 199 * 100, "arz": 49, "bt": 199, "ns": 361 * 2000}
NS_BC = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"],
         "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}
BT_P = dict(T=600, X=200, dt=1, dx=1, dosage_termination_threshold=0.1, t1_detection_threshold=0.8,
            t2_detection_threshold=0.16, alpha=0.04, alpha_beta_ratio=10, k=1e5, t1_detection_radius=15,
            t1_death_radius=35, total_dosage=61.2)
_quiet = lambda: contextlib.redirect_stdout(io.StringIO())  # noqa: E731


def kind() -> str:
    from oracle import ref_loader
    return "reference" if ref_loader.reference_available() else "port"


def rd_beta():
    return 50 * np.cos(8 * np.arccos(np.linspace(0, 1, 201)))


def transport_beta():
    return (5 * np.cos(7.35 * np.arccos(np.linspace(0, 1, 100)))).astype(np.float32)


def bt_ic(X=200, nx=201):
    xs = np.linspace(0, X, nx)
    return 0.8 * 1e5 * np.exp(-0.25 * (xs ** 2))


# ---- one env + its action source per workload; `step()` advances it by one control step incl. auto-reset --------
def _make_reference(workload, seed):
    from oracle.ref_loader import load_reference, make_shimmed_norm_reward
    ref = load_reference()
    rng = np.random.default_rng(seed)
    if workload == "rd":
        # parabolic.py:126-164 with a float64 state buffer (the fp64 mode of SURVEY 8(c): same code, u re-created as f64)
        beta = rd_beta()
        box = {}
        env = ref.ReactionDiffusionPDE1D(
            reward_class=make_shimmed_norm_reward(ref, nt=100000, norm=2, horizon="temporal"), sensing_noise_func=lambda s: s,
            reset_init_condition_func=lambda nx: box["ic"], reset_recirculation_func=lambda nx: beta, normalize=True,
            sensing_loc="full", control_type="Dirchilet", sensing_type=None, limit_pde_state_size=False, max_control_value=20,
            T=1.0, dt=1e-5, X=1.0, dx=5e-3, control_sample_rate=1e-3)

        def reset():
            box["ic"] = np.ones(201) * rng.uniform(1, 10)
            env.reset()
            env.u = np.zeros(env.u.shape, dtype=np.float64)
            env.u[0] = box["ic"]
            env.beta = np.asarray(env.beta, dtype=np.float64)
        act = lambda: np.array([rng.uniform(-1, 1)], dtype=np.float32)  # noqa: E731
    elif workload == "transport1":
        # hyperbolic.py:126-169; config 1: open-loop zero boundary control, normalize=False
        beta = transport_beta()
        box = {}
        env = ref.TransportPDE1D(
            reward_class=ref.TunedReward1D(10000, -1e3, 3e2), sensing_noise_func=lambda s: s,
            reset_init_condition_func=lambda nx: box["ic"], reset_recirculation_func=lambda nx: beta, normalize=False,
            sensing_loc="full", control_type="Dirchilet", sensing_type=None, limit_pde_state_size=False, max_control_value=20,
            T=1.0, dt=1e-4, X=1.0, dx=1e-2, control_sample_rate=0.1)

        def reset():
            box["ic"] = np.ones(100) * rng.uniform(1, 10)
            env.reset()
        act = lambda: np.zeros(1, dtype=np.float32)  # noqa: E731
    elif workload == "arz":
        with _quiet():
            env = ref.TrafficPDE1D(simulation_type="outlet", control_freq=1, reward_class=ref.TrafficARZReward(),
                                   T=240, dt=0.25, X=500, dx=10)

        def reset():
            with _quiet():
                env.reset()
        act = lambda: rng.uniform(0.8 * env.qs, 1.2 * env.qs, 1)  # noqa: E731
    elif workload == "bt":
        env = ref.BrainTumor1D(reward_class=ref.BrainTumorReward(), normalize=True, reset_init_condition_func=bt_ic,
                               verbose=False, D=0.2 * rng.uniform(0.9, 1.1), rho=0.03 * rng.uniform(0.9, 1.1), **BT_P)

        def reset():
            with _quiet():
                env.reset()
        act = lambda: rng.uniform(0, 0.1)  # noqa: E731
    elif workload == "ns":
        box = {}
        env = ref.NavierStokes2D(reset_init_condition_func=lambda X: tuple(np.full((21, 21), c) for c in box["c"]),
                                 boundary_condition=NS_BC, U_ref=np.zeros((200, 21, 21, 2)), action_ref=2.0 * np.ones(1000),
                                 reward_class=ref.NSReward(0.1), action_dim=1, T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05)

        def reset():
            box["c"] = rng.uniform(-5, 5, 3)
            env.reset()
        act = lambda: rng.uniform(2, 4)  # noqa: E731
    else:
        raise KeyError(workload)
    reset()

    def step():
        with _quiet():
            out = env.step(act())
        if out is None or bool(out[2]) or bool(out[3]):
            reset()
    return step


def _make_port(workload, seed):
    from oracle import np_oracle as O
    rng = np.random.default_rng(seed)
    if workload == "rd":
        e = O.ReactionDiffusionOracle(1, reward=O.NormRewardSpec(100000, norm=2), precision="fp64", T=1.0, dt=1e-5, X=1.0, dx=5e-3,
                                      control_sample_rate=1e-3, normalize=True, max_control_value=20)
        reset = lambda: e.reset(np.ones(201) * rng.uniform(1, 10), rd_beta())  # noqa: E731
        act = lambda: np.array([rng.uniform(-1, 1)], dtype=np.float32)  # noqa: E731
    elif workload == "transport1":
        e = O.TransportOracle(1, reward=O.TunedReward1DSpec(10000, -1e3, 3e2), T=1.0, dt=1e-4, X=1.0, dx=1e-2, control_sample_rate=0.1,
                              normalize=False, max_control_value=20)
        reset = lambda: e.reset(np.ones(100) * rng.uniform(1, 10), transport_beta())  # noqa: E731
        act = lambda: np.zeros(1, dtype=np.float32)  # noqa: E731
    elif workload == "arz":
        e = O.TrafficARZOracle(1, simulation_type="outlet", control_freq=1, T=240, dt=0.25, X=500, dx=10)
        reset = lambda: e.reset()  # noqa: E731
        act = lambda: rng.uniform(0.96, 1.44, (1, 1))  # noqa: E731
    elif workload == "bt":
        e = O.BrainTumorOracle(1, D=0.2 * rng.uniform(0.9, 1.1), rho=0.03 * rng.uniform(0.9, 1.1), **BT_P)
        reset = lambda: e.reset(bt_ic())  # noqa: E731
        act = lambda: rng.uniform(0, 0.1, 1)  # noqa: E731
    elif workload == "ns":
        e = O.NavierStokesOracle(1, boundary_condition=NS_BC, U_ref=np.zeros((200, 21, 21, 2)), action_ref=2.0 * np.ones(1000),
                                 T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05)
        reset = lambda: e.reset(*(np.full((21, 21), rng.uniform(-5, 5)) for _ in range(3)))  # noqa: E731
        act = lambda: rng.uniform(2, 4, 1)  # noqa: E731
    else:
        raise KeyError(workload)
    reset()

    def step():
        out = e.step(act())
        if bool(np.any(out[2])) or bool(np.any(out[3])) or (len(out) > 4 and not bool(np.all(out[4]))):
            reset()
    return step


def make_env(workload, seed, impl=None):
    impl = impl or kind()
    return _make_reference(workload, seed) if impl == "reference" else _make_port(workload, seed)


def _worker(conn, workload, k, seed, impl):
    warnings.simplefilter("ignore")
    np.seterr(all="ignore")
    steps = [make_env(workload, seed * 1000 + i, impl) for i in range(k)]
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg is None:
            break
        for _ in range(msg):                # msg lockstep rounds per message (1 = strict AsyncVectorEnv lockstep)
            for s in steps:
                s()
        conn.send(msg)


def lockstep_rate(workload, cores=None, k_envs=1, rounds=None, seconds=None, warmup=2, impl=None):
    """env-steps/s of `cores` worker processes x k_envs envs advanced in lockstep through pipes.  Either a fixed number of
    `rounds` or a time budget in `seconds` (then the round count is whatever fits).  -> dict(value, rounds, wall_s, cores,
    kind, envs, sample)."""
    import multiprocessing as mp
    impl = impl or kind()
    cores = cores or len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    conns, procs = [], []
    for w in range(cores):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_worker, args=(b, workload, k_envs, 17 + w, impl), daemon=True)
        p.start()
        conns.append(a)
        procs.append(p)
    for c in conns:
        assert c.recv() == "ready"

    def one():
        for c in conns:
            c.send(1)
        for c in conns:
            c.recv()

    for _ in range(warmup):
        one()
    n, t0 = 0, time.perf_counter()
    while (rounds is not None and n < rounds) or (rounds is None and time.perf_counter() - t0 < seconds):
        one()
        n += 1
    wall = time.perf_counter() - t0
    for c in conns:
        c.send(None)
    for p in procs:
        p.join(timeout=5)
        if p.is_alive():            # never leave a worker behind: the GPU legs that follow are timed on the same host
            p.terminate()
            p.join(timeout=5)
    what = ("the reference's own classes (unmodified files staged in oracle/_ref, numpy)" if impl == "reference"
            else "numpy oracle (port of the reference's numpy path; no reference tree found)")
    return {"value": n * cores * k_envs / wall, "rounds": n, "wall_s": wall, "cores": cores, "kind": impl, "envs": cores * k_envs,
            "sample": f"{what}: {cores} worker process(es) x {k_envs} env(s) in lockstep over pipes (AsyncVectorEnv-equivalent, "
                      f"same-step auto-reset), {n} control steps each, {wall:.1f} s wall"}

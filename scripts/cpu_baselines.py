"""Host-CPU baselines for every BASELINE.json configuration: the numpy oracle (a port of the reference's numpy path;
oracle/np_oracle.py) stepped like gymnasium's AsyncVectorEnv -- one worker process per host core, `k` single-env
instances per worker, lock-stepped through pipes -- on the box this script runs on.  bench.py does the same for the
headline configuration; this script covers the others so that scripts/bench_all_envs.py's GPU numbers have a CPU number
from the same box beside them.  Prints one JSON line per configuration.

    python scripts/cpu_baselines.py [--seconds 4]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def make_env(name, seed):
    """-> (step() closure advancing ONE env by one control step incl. auto-reset, grid-point updates per env-step)"""
    from oracle import np_oracle as O
    rng = np.random.default_rng(seed)
    if name == "transport":
        kw = dict(T=1.0, dt=1e-4, X=1.0, dx=1e-2, control_sample_rate=0.1, normalize=True, max_control_value=20)
        beta = (5 * np.cos(7.35 * np.arccos(np.linspace(0, 1, 100)))).astype(np.float32)
        e = O.TransportOracle(1, reward=O.TunedReward1DSpec(10000, -1e3, 3e2), **kw)
        reset = lambda: e.reset(np.ones(100) * rng.uniform(1, 10), beta)  # noqa: E731
        act = lambda: np.array([rng.uniform(-1, 1)], dtype=np.float32)  # noqa: E731
        pts = 99 * 1000
    elif name == "rd_ref":
        kw = dict(T=1.0, dt=1e-5, X=1.0, dx=5e-3, control_sample_rate=1e-3, normalize=True, max_control_value=20)
        beta = (50 * np.cos(8 * np.arccos(np.linspace(0, 1, 201)))).astype(np.float32)
        e = O.ReactionDiffusionOracle(1, reward=O.TunedReward1DSpec(100000, -1e3, 3e2), **kw)
        reset = lambda: e.reset(np.ones(201) * rng.uniform(1, 10), beta)  # noqa: E731
        act = lambda: np.array([rng.uniform(-1, 1)], dtype=np.float32)  # noqa: E731
        pts = 199 * 100
    elif name == "arz":
        e = O.TrafficARZOracle(1, simulation_type="outlet", control_freq=1, T=240, dt=0.25, X=500, dx=10)
        reset = lambda: e.reset()  # noqa: E731
        act = lambda: rng.uniform(0.96, 1.44, (1, 1))  # noqa: E731
        pts = 49
    elif name == "brain_tumor":
        xs = np.linspace(0, 200, 201)
        ic = 0.8 * 1e5 * np.exp(-0.25 * xs ** 2)
        e = O.BrainTumorOracle(1, total_dosage=61.2, T=600, dt=1, X=200, dx=1)
        reset = lambda: e.reset(ic)  # noqa: E731
        act = lambda: rng.uniform(0, 0.1, 1)  # noqa: E731
        pts = 199
    elif name == "navier_stokes":
        bc = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"], "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}
        e = O.NavierStokesOracle(1, boundary_condition=bc, U_ref=np.zeros((200, 21, 21, 2)), action_ref=2.0 * np.ones(1000),
                                 T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05)
        reset = lambda: e.reset(*(np.full((21, 21), rng.uniform(-5, 5)) for _ in range(3)))  # noqa: E731
        act = lambda: rng.uniform(2, 4, 1)  # noqa: E731
        pts = 361 * 2000
    else:
        raise KeyError(name)
    reset()

    def step():
        out = e.step(act())
        te, tr = out[2], out[3]
        if bool(np.any(te)) or bool(np.any(tr)):
            reset()
    return step, pts


def worker(conn, name, k, seed):
    steps = [make_env(name, seed * 100 + i)[0] for i in range(k)]
    conn.send("ready")
    while conn.recv() is not None:
        for s in steps:
            s()
        conn.send(1)


def lockstep_rate(name, seconds, cores, k):
    ctx = mp.get_context("fork")
    conns, procs = [], []
    for w in range(cores):
        a, b = ctx.Pipe()
        p = ctx.Process(target=worker, args=(b, name, k, w), daemon=True)
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
    one()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        one()
        n += 1
    dt = time.perf_counter() - t0
    for c in conns:
        c.send(None)
    for p in procs:
        p.join(timeout=5)
    return n * cores * k / dt, n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    cores = len(os.sched_getaffinity(0))
    for name in ("transport", "rd_ref", "arz", "brain_tumor", "navier_stokes"):
        if args.only and name not in args.only.split(","):
            continue
        k = 1 if name in ("transport", "navier_stokes") else 4
        pts = make_env(name, 0)[1]
        rate, rounds = lockstep_rate(name, args.seconds, cores, k)
        print(json.dumps({"config": name, "impl": "numpy oracle (port of the reference's numpy path), AsyncVectorEnv-equivalent",
                          "cores": cores, "envs": cores * k, "lockstep_rounds": rounds, "env_steps_per_sec": rate,
                          "grid_pt_updates_per_sec": rate * pts}), flush=True)


if __name__ == "__main__":
    main()

"""Device-resident throughput of every env family on one B200 (the BASELINE.json configs other than the
headline one, which bench.py measures).  Prints one JSON line per configuration.

    python scripts/bench_all_envs.py [--quick]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pdecontrolgym_b200 as pcg

dev = torch.device("cuda:0")
QUICK = "--quick" in sys.argv


def timed(step_fn, steps, warmup=5):
    for i in range(warmup):
        step_fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        step_fn(warmup + i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def report(name, n_envs, ms, pts_per_env_step, extra=None):
    rate = n_envs / (ms * 1e-3)
    line = {"config": name, "n_envs": n_envs, "ms_per_batched_step": ms, "env_steps_per_sec": rate,
            "grid_pt_updates_per_sec": rate * pts_per_env_step}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def bench_bc1d(env_id, name, N, precision, kw, npts, beta, reward, pts, steps=30, beta_dtype=np.float32):
    env = pcg.make(env_id, num_envs=N, device=dev, precision=precision, reward_class=reward, autoreset=True, copy_obs=False, **kw)
    ic = torch.empty(N, 1, dtype=torch.float64, device=dev).uniform_(1, 10) * torch.ones(1, npts, dtype=torch.float64, device=dev)
    env.reset(options={"init_condition": ic, "beta": beta.astype(beta_dtype)})
    acts = torch.empty(steps + 5, N, device=dev).uniform_(-1, 1)
    ms = timed(lambda i: env.step(acts[i]), steps)
    report(name, N, ms, pts, {"precision": precision})
    del env


def bench_transport():
    # ---- Transport: config 1 (single env, latency) and the PPO-rollout shape (2^18 envs per GPU) ----
    tkw = dict(T=1.0, dt=1e-4, X=1.0, dx=1e-2, control_sample_rate=0.1, normalize=True, max_control_value=20)
    tbeta = 5 * np.cos(7.35 * np.arccos(np.linspace(0, 1, 100)))
    for N in ([1, 65536] if QUICK else [1, 4096, 65536, 262144]):
        bench_bc1d("PDEControlGym-TransportPDE1D", f"TransportPDE1D ref(f32) x{N}", N, "ref", tkw, 100, tbeta,
                   pcg.TunedReward1D(10000, -1e3, 3e2), 99 * 1000, steps=8 if N > 1000 else 20)
    bench_bc1d("PDEControlGym-TransportPDE1D", "TransportPDE1D fp64 x65536", 65536, "fp64", tkw, 100, tbeta,
               pcg.TunedReward1D(10000, -1e3, 3e2), 99 * 1000, steps=6, beta_dtype=np.float64)


def bench_rd():
    # ---- ReactionDiffusion in the reference's own precision (f32 state) ----
    rkw = dict(T=1.0, dt=1e-5, X=1.0, dx=5e-3, control_sample_rate=1e-3, normalize=True, max_control_value=20)
    rbeta = 50 * np.cos(8 * np.arccos(np.linspace(0, 1, 201)))
    bench_bc1d("PDEControlGym-ReactionDiffusionPDE1D", "ReactionDiffusionPDE1D ref(f32, beta f32) x65536", 65536, "ref", rkw, 201, rbeta,
               pcg.TunedReward1D(100000, -1e3, 3e2), 199 * 100)
    bench_bc1d("PDEControlGym-ReactionDiffusionPDE1D", "ReactionDiffusionPDE1D fp32-fast x65536", 65536, "fp32", rkw, 201, rbeta,
               pcg.TunedReward1D(100000, -1e3, 3e2), 199 * 100)
    bench_bc1d("PDEControlGym-ReactionDiffusionPDE1D", "ReactionDiffusionPDE1D fp64 x65536 (TunedReward1D)", 65536, "fp64", rkw, 201, rbeta,
               pcg.TunedReward1D(100000, -1e3, 3e2), 199 * 100, beta_dtype=np.float64)


def bench_arz():
    # ---- Traffic ARZ x16,384 (config 3) ----
    N = 16384
    env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=dev, simulation_type="outlet", reward_class=pcg.TrafficARZReward(),
                   T=240, dt=0.25, X=500, dx=10, copy_obs=False)
    env.reset()
    acts = torch.empty(64, N, 1, dtype=torch.float64, device=dev).uniform_(0.96, 1.44)
    ms = timed(lambda i: env.step(acts[i % 64]), 200)
    report("TrafficPDE1D outlet x16384 (control_freq=1)", N, ms, 49, {"hbm_gbs": N * (2 * 51 * 8 * 2 + 102 * 8 + 40) / (ms * 1e-3) / 1e9})
    g = pcg.GraphedStep(env, acts[0])
    env.reset()
    ms = timed(lambda i: g.step(acts[i % 64]), 200)
    report("TrafficPDE1D outlet x16384 (control_freq=1), CUDA-graph replay", N, ms, 49, {"hbm_gbs": N * (2 * 51 * 8 * 2 + 102 * 8 + 40) / (ms * 1e-3) / 1e9})
    del env, g
    # rollout-style stepping: private observation copies + same-step auto-reset, all inside the one step launch
    env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=dev, simulation_type="outlet", reward_class=pcg.TrafficARZReward(),
                   T=240, dt=0.25, X=500, dx=10, copy_obs=True, autoreset=True)
    env.reset()
    ms = timed(lambda i: env.step(acts[i % 64]), 200)
    report("TrafficPDE1D outlet x16384 (control_freq=1), autoreset + observation copy (1 launch/step)", N, ms, 49)
    del env
    N = 262144
    env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=dev, simulation_type="outlet", reward_class=pcg.TrafficARZReward(),
                   T=240, dt=0.25, X=500, dx=10, copy_obs=False)
    env.reset()
    acts = torch.empty(8, N, 1, dtype=torch.float64, device=dev).uniform_(0.96, 1.44)
    ms = timed(lambda i: env.step(acts[i % 8]), 100)
    report("TrafficPDE1D outlet x262144 (control_freq=1)", N, ms, 49, {"hbm_gbs": N * (2 * 51 * 8 * 2 + 102 * 8 + 40) / (ms * 1e-3) / 1e9})
    del env


def bench_bt():
    # ---- BrainTumor x16,384 (config 3) ----
    def bt_ic(X, nx, n=None):
        xs = np.linspace(0, X, nx)
        return 0.8 * 1e5 * np.exp(-0.25 * (xs ** 2))
    for N in (16384, 262144):
        rng = np.random.default_rng(0)
        env = pcg.make("PDEControlGym-BrainTumor1D", num_envs=N, device=dev, reward_class=pcg.BrainTumorReward(), total_dosage=61.2,
                       D=0.2 * rng.uniform(0.9, 1.1, N), rho=0.03 * rng.uniform(0.9, 1.1, N), T=600, dt=1, X=200, dx=1, copy_obs=False)
        env.reset(options={"init_condition": bt_ic(200, 201)})
        acts = torch.empty(16, N, dtype=torch.float64, device=dev).uniform_(0, 0.1)
        ms = timed(lambda i: env.step(acts[i % 16]), 150)
        report(f"BrainTumor1D raw step x{N}", N, ms, 199, {"hbm_gbs": N * (2 * 201 * 8 + 64 + 32 + 12) / (ms * 1e-3) / 1e9})
        if N == 16384:
            env.reset(options={"init_condition": bt_ic(200, 201)})
            env.autoreset = True        # same-step auto-reset fused into the step launch
            ms = timed(lambda i: env.step(acts[i % 16]), 150)
            report(f"BrainTumor1D raw step x{N}, autoreset (1 launch/step)", N, ms, 199)
            env.autoreset = False
            env.reset(options={"init_condition": bt_ic(200, 201)})
            g = pcg.GraphedStep(env, acts[0])
            env.reset(options={"init_condition": bt_ic(200, 201)})
            ms = timed(lambda i: g.step(acts[i % 16]), 150)
            report(f"BrainTumor1D raw step x{N}, CUDA-graph replay", N, ms, 199, {"hbm_gbs": N * (2 * 201 * 8 + 64 + 32 + 12) / (ms * 1e-3) / 1e9})
            del g
            env.reset(options={"init_condition": bt_ic(200, 201)})
        w = pcg.TherapyWrapper(env)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        w.benchmark()
        w.reset()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        days = float(env.t_benchmark.double().mean() + env.growthDays.double().mean())
        report(f"BrainTumor1D TherapyWrapper.benchmark()+reset() x{N} (~{days:.0f} days/env in 2 launches)", N, dt * 1e3 / days, 199)
        del env, w


def bench_ns():
    # ---- NavierStokes2D x1,024 (config 4) ----
    bc = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"], "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}
    for N, prec in ([(1024, "ref")] if QUICK else [(1024, "ref"), (8192, "ref"), (1024, "fp32"), (8192, "fp32")]):
        uref = np.zeros((200, 21, 21, 2))
        env = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=dev, boundary_condition=bc, U_ref=uref, action_ref=2.0 * np.ones(1000),
                       reward_class=pcg.NSReward(0.1), action_dim=1, T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05, copy_obs=False, precision=prec)
        c = torch.empty(3, N, 1, 1, dtype=torch.float64, device=dev).uniform_(-5, 5) * torch.ones(1, 21, 21, dtype=torch.float64, device=dev)
        env.reset(options={"init": (c[0], c[1], c[2])})
        acts = torch.empty(32, N, dtype=torch.float64, device=dev).uniform_(2, 4)
        ms = timed(lambda i: env.step(acts[i % 32]), 20, warmup=3)
        report(f"NavierStokes2D 21x21, 2000 Jacobi sweeps x{N}" + (" (opt-in float32 mode)" if prec == "fp32" else ""), N, ms, 361 * 2000)
        del env


def bench_large():
    # ---- grids beyond the reference's: CTA-per-env shared-memory kernels ----
    for nx, N in ((1000, 16384), (4000, 4096)):
        dx = 1.0 / nx
        rkw = dict(T=1.0, dt=0.4 * dx * dx, X=1.0, dx=dx, control_sample_rate=100 * 0.4 * dx * dx, normalize=True, max_control_value=20)
        rbeta = 50 * np.cos(8 * np.arccos(np.linspace(0, 1, nx + 1)))
        bench_bc1d("PDEControlGym-ReactionDiffusionPDE1D", f"ReactionDiffusionPDE1D fp64, {nx + 1} points, 100 substeps x{N}", N, "fp64", rkw, nx + 1, rbeta,
                   pcg.NormReward(int(round(1.0 / rkw["dt"]))), (nx - 1) * 100, steps=10, beta_dtype=np.float64)
        tkw = dict(T=1.0, dt=dx / 100, X=1.0, dx=dx, control_sample_rate=10 * dx, normalize=True, max_control_value=20)
        tbeta = 5 * np.cos(7.35 * np.arccos(np.linspace(0, 1, nx)))
        bench_bc1d("PDEControlGym-TransportPDE1D", f"TransportPDE1D ref(f32), {nx} points, 1000 substeps x{N}", N, "ref", tkw, nx, tbeta,
                   pcg.TunedReward1D(int(round(1.0 / tkw["dt"])), -1e3, 3e2), (nx - 1) * 1000, steps=4)
    bc = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"], "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}
    for n, N in ((41, 1024), (64, 592)):
        h = 1.0 / (n - 1)
        env = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=dev, boundary_condition=bc, U_ref=np.zeros((200, n, n, 2)), action_ref=2.0 * np.ones(1000),
                       reward_class=pcg.NSReward(0.1), action_dim=1, T=150 * 1e-3 * (h / 0.05) ** 2, dt=1e-3 * (h / 0.05) ** 2, X=1, dx=h, Y=1, dy=h, copy_obs=False)
        c = torch.empty(3, N, 1, 1, dtype=torch.float64, device=dev).uniform_(-5, 5) * torch.ones(1, n, n, dtype=torch.float64, device=dev)
        env.reset(options={"init": (c[0], c[1], c[2])})
        acts = torch.empty(8, N, dtype=torch.float64, device=dev).uniform_(2, 4)
        ms = timed(lambda i: env.step(acts[i % 8]), 5, warmup=2)
        report(f"NavierStokes2D {n}x{n}, 2000 Jacobi sweeps x{N}", N, ms, (n - 2) ** 2 * 2000)
        del env


def main():
    """All families, or only those named in PDECG_BENCH_ONLY (comma list of: transport, rd, arz, bt, ns, large)."""
    only = [t for t in os.environ.get("PDECG_BENCH_ONLY", "").split(",") if t]
    for name, fn in (("transport", bench_transport), ("rd", bench_rd), ("arz", bench_arz), ("bt", bench_bt), ("ns", bench_ns), ("large", bench_large)):
        if not only or name in only:
            fn()


if __name__ == "__main__":
    main()

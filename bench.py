#!/usr/bin/env python
"""bench.py -- benchmark of the batched PDE-env hot path (one JSON line on stdout).

Headline workload (BASELINE.json configs[1]): ReactionDiffusionPDE1D x 65,536 envs per GPU, reference default
grid (T=1, dt=1e-5, dx=5e-3 -> 201 points, 100 substeps per control step), Dirichlet boundary control,
NormReward, fresh random actions every step, fp64 state and arithmetic.  A "step" is ONE batched control step
of all envs (= 65,536 env-steps = 1.304e9 grid-point updates per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E] [--no-secondary]

N > 1: launched by torchrun, one rank per GPU, env batch sharded (weak scaling: E envs per GPU), no collective
inside the step; one NCCL all-gather of per-rank episode statistics per timed region.

The JSON line also carries `secondary`: the other BASELINE.json configurations in the same schema
(value / roofline / cpu_baseline / e2e / clocks each, measured after the headline's timed region) --
config 1 (TransportPDE1D x 1 env, latency), config 3 (TrafficPDE1D and BrainTumor1D x 16,384), config 4
(NavierStokes2D x 1,024) at N = 1, and config 5 (PPO rollout collection, 2^20 envs sharded over the N GPUs) at
every N.

`--impl reference`: the reference's own CPU implementation of the headline path -- the UNMODIFIED env / reward
classes of lukebhan/PDEControlGym, staged verbatim into the git-ignored oracle/_ref/ by __graft_entry__.build()
(oracle/make_ref.py) and stepped in lockstep by an AsyncVectorEnv-equivalent (one worker process per host core).
`cpu_baseline.kind` says "reference"; only if no reference tree can be found does the numpy oracle stand in ("port").
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# ---- the headline workload ------------------------------------------------------------------------
NX_PTS = 201
SUBSTEPS = 100
INTERIOR = 199                       # points updated per substep (parabolic.py:143: u[1:Nx])
FLOPS_PER_POINT = 7                  # sub, mul(2*u), add, mul(F), add, mul(c*u), add -- dt*beta hoisted
FP64_INSTR_PER_POINT = 6             # what the kernel issues: (p_l - 2p) is one exact FMA
RD_KW = dict(T=1.0, dt=1e-5, X=1.0, dx=5e-3, control_sample_rate=1e-3, normalize=True,
             max_control_value=20, sensing_loc="full", control_type="Dirchilet", sensing_type=None,
             limit_pde_state_size=False)
REWARD_NT = 100000
NS_BC = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"],
         "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}


def workload_name(envs):
    return (f"ReactionDiffusionPDE1D x{envs} envs/GPU, 201 pts, 100 substeps/step, Dirichlet, NormReward(L2,temporal), "
            "random actions, fp64")


def headline_config(envs, world):
    """The `config` object of BOTH arms (the reference arm times a bounded sample of this same workload)."""
    return {"workload": workload_name(envs), "envs_per_gpu": envs, "global_envs": world * envs,
            "parallelism": f"env-shard x{world} (no collective in step; 1 stats all-gather per timed region)",
            "l2": "no flush: state (105 MB) + observation copy (105 MB) written every step exceed the 126 MB L2",
            "autoreset": True}


def beta_fn():
    return 50 * np.cos(8 * np.arccos(np.linspace(0, 1, NX_PTS)))


# ---- clocks sampler ------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed regions (NVML in a thread, every 2 ms; the
    `nvidia-smi -lms` loop of the profiling recipe is only the fallback because each of its queries stalls kernel
    launches for milliseconds and distorts a sub-second timed region).  `window(t0, t1)` summarises the samples
    taken inside one timed region; one sampler serves the headline and every secondary configuration."""
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period=0.002):
        self.index, self.period, self.rows, self.stop_flag = index, period, [], False
        self.thread, self.h, self.nv, self.proc = None, None, None, None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.nv = None
            try:
                q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                              "--format=csv,noheader,nounits", "-lms", "500"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._pump, daemon=True)
                self.thread.start()
            except Exception:  # noqa: BLE001
                self.proc = None
        return self

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.time(), mhz, pw, bits))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def _pump(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                bits = sum(self.BITS[n] for n, v in zip(names, f[3:7]) if v.lower().startswith("active"))
                self.max_mhz = float(f[1])
                self.rows.append((time.time(), float(f[0]), float(f[2]), bits))
            except (ValueError, IndexError):
                continue

    def window(self, t0, t1):
        if self.proc is not None:
            time.sleep(0.6)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or [r for r in self.rows if t0 - 0.2 <= r[0] <= t1 + 0.2]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}
        bits = 0
        for r in rows:
            bits |= r[3]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": getattr(self, "max_mhz", None),
                "power_w_max": max(r[2] for r in rows), "samples": len(rows),
                "reasons": sorted(n for n, b in self.BITS.items() if bits & b),
                "source": "nvml" if self.nv is not None else "nvidia-smi"}

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        if self.thread is not None:
            self.thread.join(timeout=1.0)


# ---- the reference arm -------------------------------------------------------------------------------
def run_reference_arm(args):
    """The reference's own CPU path (oracle/cpu_arm.py) on the headline config; each bench step = a bounded sample of
    `ROUNDS` lockstep control steps of (host cores x 1) reference envs -- one env per worker process, like
    gymnasium.vector.AsyncVectorEnv."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_arm
    rounds_per_step = 25
    res = cpu_arm.lockstep_rate("rd", k_envs=1, rounds=args.steps * rounds_per_step, warmup=args.warmup * rounds_per_step)
    rate = res["value"]
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": rate, "unit": "env-steps/s",
        "grid_pt_updates_per_sec": rate * INTERIOR * SUBSTEPS,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * res["wall_s"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": headline_config(args.envs, args.gpus),
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"] + f"; one bench step = {rounds_per_step} lockstep control steps of {res['envs']} envs"},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- helpers of our arm --------------------------------------------------------------------------------
def _pinned(torch, *shape, dtype):
    return torch.empty(*shape, dtype=dtype).pin_memory()


def _event_loop(torch, dev, fn, steps, warmup):
    """CUDA-event time of `steps` calls of fn(i) on the current stream after `warmup` calls -> (ms per call, t0, t1)."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize(dev)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    s.record()
    for i in range(steps):
        fn(warmup + i)
    e.record()
    torch.cuda.synchronize(dev)
    return s.elapsed_time(e) / steps, t0, time.time()


def _host_loop(torch, dev, fn, steps, warmup):
    """End-to-end time of `steps` host-buffer calls (each synchronises): max(CUDA events, wall clock) per call."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize(dev)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    s.record()
    for i in range(steps):
        fn(warmup + i)
    e.record()
    torch.cuda.synchronize(dev)
    w1 = time.perf_counter()
    return max(s.elapsed_time(e), 1e3 * (w1 - w0)) / steps


def _kernel_ms(torch, dev, fn, steps, warmup=3):
    """CUDA-event time of ONE call of fn (a single kernel launch, nothing else on the stream), events on the launching
    stream: the mean over `steps` individually bracketed launches, or -- for kernels so short that the two event records
    per launch outweigh them -- the average over one back-to-back run of the same launches, whichever is smaller."""
    for i in range(warmup):
        fn(i)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize(dev)
    for i in range(steps):
        ev[i][0].record()
        fn(warmup + i)
        ev[i][1].record()
    torch.cuda.synchronize(dev)
    each = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        fn(warmup + i)
    e.record()
    torch.cuda.synchronize(dev)
    return min(each, s.elapsed_time(e) / steps)


def _probe_peak(torch, dev, L, _lib, dtype):
    ms, ops = C.c_float(), C.c_double()
    _lib.check(L.pdecg_probe_fp_issue(dtype, 148 * 8, 256, 20000, C.byref(ms), C.byref(ops), torch.cuda.current_stream(dev).cuda_stream))
    return ops.value / (ms.value * 1e-3) / 1e12


def _cpu(workload, seconds, cores=None, k_envs=1):
    from oracle import cpu_arm
    r = cpu_arm.lockstep_rate(workload, cores=cores, k_envs=k_envs, seconds=seconds, warmup=2)
    return {"value": r["value"], "unit": "env-steps/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}


# ---- secondary configurations (BASELINE.json configs 1, 3, 4 at N = 1; config 5 at every N) ---------------
def secondary_transport1(torch, pcg, dev, sampler, peaks, cpu_seconds):
    """Config 1: TransportPDE1D, ONE env, T=1, dt=1e-4, dx=1e-2, 1000 substeps per control step, open-loop zero control,
    TunedReward1D -- a latency measurement: one warp walks a chain of 1000 dependent stencil updates per step."""
    from oracle.cpu_arm import transport_beta
    kw = dict(T=1.0, dt=1e-4, X=1.0, dx=1e-2, control_sample_rate=0.1, normalize=False, max_control_value=20)
    mk = lambda **x: pcg.make("PDEControlGym-TransportPDE1D", num_envs=1, device=dev, precision="ref",  # noqa: E731
                              reward_class=pcg.TunedReward1D(10000, -1e3, 3e2), autoreset=True, **x, **kw)
    opts = {"init_condition": np.ones(100) * 5.5, "beta": transport_beta()}
    env = mk()
    env.reset(options=opts)
    a = torch.zeros(1, dtype=torch.float32, device=dev)
    ms, t0, t1 = _event_loop(torch, dev, lambda i: env.step(a), 200, 10)
    envk = mk(copy_obs=False)
    envk.reset(options=opts)
    k_ms = _kernel_ms(torch, dev, lambda i: envk.step(a), 100)
    # the same step replayed from a CUDA graph (pcg.GraphedStep): no Python / ctypes between the launches
    envg = mk()
    envg.reset(options=opts)
    gs = pcg.GraphedStep(envg, a)
    envg.reset(options=opts)
    g_ms, _, _ = _event_loop(torch, dev, lambda i: gs.step(a), 200, 10)
    envh = mk()
    envh.reset(options=opts)
    ah, oh = _pinned(torch, 1, dtype=torch.float32).zero_(), _pinned(torch, 1, 100, dtype=torch.float32)
    rh, teh, trh = _pinned(torch, 1, dtype=torch.float64), _pinned(torch, 1, dtype=torch.uint8), _pinned(torch, 1, dtype=torch.uint8)
    e2e_ms = _host_loop(torch, dev, lambda i: envh.step_host(ah, oh, rh, teh, trh), 200, 10)
    # dependent chain per substep on the critical lane: SHFL (neighbour's new value, ~24 cycles) + 7 dependent FP32 ops
    # of 4 cycles each (sub, the 3-op exact quotient, add, mul, add; B300_MICROARCH.md: fma/alu latency 4)
    chain_cycles, clk = 24 + 7 * 4, (peaks.get("sm_max_mhz") or 1965.0) * 1e6
    peak = clk / chain_cycles
    achieved = 1000.0 / (k_ms * 1e-3)
    return {
        "config": "configs[0]", "workload": "TransportPDE1D x1 env, 100 pts, 1000 substeps/step (T=1, dt=1e-4, dx=1e-2), zero control, TunedReward1D, ref precision (f32)",
        "metric": "env_steps_per_sec", "value": 1.0 / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms, "n_envs": 1,
        "grid_pt_updates_per_sec": 99 * 1000 / (ms * 1e-3), "dtype": "f32", "gpu_launches_per_step": 1,
        "graph_replay": {"ms_per_step": g_ms, "value": 1.0 / (g_ms * 1e-3), "api": "pcg.GraphedStep(env, action).step(action)"},
        "roofline": {"bound": "latency", "achieved": achieved, "peak": peak, "unit": "dependent substeps/s", "frac": achieved / peak,
                     "kernel": "bc1d_step_kernel<Transport, ModeRefB32, K=4, fast, 32 lanes/env>", "kernel_ms": k_ms, "traffic": None,
                     "peak_source": f"one warp, 1000 dependent substeps per launch: modelled critical path {chain_cycles} cycles per substep "
                                    "(SHFL ~24 + 7 dependent FP32 ops x 4) at the max SM clock; no throughput roof applies to a single env"},
        "cpu_baseline": _cpu("transport1", cpu_seconds, cores=1, k_envs=1),
        "e2e": {"value": 1.0 / (e2e_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": 4,
                "d2h_bytes_per_step": 100 * 4 + 8 + 2, "api": "env.step_host -> pdecg_bc1d_step_host (pinned host buffers)"},
        "clocks": sampler.window(t0, t1),
    }


def _graph_ms(torch, dev, pcg, env, acts, reset, steps=300, warmup=20):
    """The env's step replayed from a CUDA graph (pcg.GraphedStep, the repo's API for launch-bound batches: one
    cudaGraphLaunch per step, no Python / ctypes between the kernels).  Returns (ms per `g.step(action)` -- action copy +
    replay, CUDA events --, ms per bare replay back to back, t0, t1)."""
    gs = pcg.GraphedStep(env, acts[0])
    reset()                                   # the capture advanced the state
    ms, t0, t1 = _event_loop(torch, dev, lambda i: gs.step(acts[i % acts.shape[0]]), steps, warmup)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    s.record()
    for _ in range(steps):
        gs.graph.replay()
    e.record()
    torch.cuda.synchronize(dev)
    return ms, s.elapsed_time(e) / steps, t0, t1


def secondary_arz(torch, pcg, dev, sampler, peaks, cpu_seconds, N=16384):
    """Config 3a: TrafficPDE1D (ARZ) x 16,384, tutorial defaults, outlet control, control_freq=1."""
    kw = dict(T=240, dt=0.25, X=500, dx=10, simulation_type="outlet", reward_class=pcg.TrafficARZReward())
    env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=dev, autoreset=True, **kw)
    env.reset()
    acts = torch.empty(64, N, 1, dtype=torch.float64, device=dev).uniform_(0.96, 1.44)
    eager_ms, _, _ = _event_loop(torch, dev, lambda i: env.step(acts[i % 64]), 300, 20)
    # a ~16 us kernel: the eager step is bound by the box's host CPU (Python + ctypes per launch), so the step is also
    # replayed from a CUDA graph -- the number a rollout loop gets whatever the host
    envg = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=dev, autoreset=True, **kw)
    envg.reset()
    ms, _, t0, t1 = _graph_ms(torch, dev, pcg, envg, acts, envg.reset)
    envk = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=dev, copy_obs=False, **kw)
    envk.reset()
    k_ms = _kernel_ms(torch, dev, lambda i: envk.step(acts[i % 64]), 200)
    envk.reset()
    k_ms = min(k_ms, _graph_ms(torch, dev, pcg, envk, acts, envk.reset, steps=200)[1])
    envh = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=dev, autoreset=True, **kw)
    envh.reset()
    ah = _pinned(torch, 64, N, 1, dtype=torch.float64).uniform_(0.96, 1.44)
    oh, rh = _pinned(torch, N, 102, dtype=torch.float64), _pinned(torch, N, dtype=torch.float64)
    teh, trh = _pinned(torch, N, dtype=torch.uint8), _pinned(torch, N, dtype=torch.uint8)
    e2e_ms = _host_loop(torch, dev, lambda i: envh.step_host(ah[i % 64], oh, rh, teh, trh), 100, 5)
    best_ms = min(ms, eager_ms)
    bytes_per_env = 8 * (2 * 51 + 2 * 51 + 102) + 8 + 6 * 8 + 8 + 8 + 2      # rows in + out, observation, action, fenv, time, reward, flags
    gbs = N * bytes_per_env / (k_ms * 1e-3) / 1e9
    return {
        "config": "configs[2] (TrafficPDE1D)", "workload": f"TrafficPDE1D (ARZ) x{N} envs, M=51, outlet control, control_freq=1, tutorial defaults, f64",
        "metric": "env_steps_per_sec", "value": N / (best_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": best_ms, "n_envs": N,
        "api": "the faster of the two public stepping APIs on this box (autoreset=True, private observation copies): " + ("graph_replay" if ms <= eager_ms else "eager_step"),
        "eager_step": {"ms_per_step": eager_ms, "value": N / (eager_ms * 1e-3), "api": "env.step(action): one Python + ctypes dispatch per step (host-bound at this batch: moves with the box's CPU)"},
        "graph_replay": {"ms_per_step": ms, "value": N / (ms * 1e-3), "api": "pcg.GraphedStep(env, action).step(action): action copy + CUDA-graph replay of env.step"},
        "grid_pt_updates_per_sec": N * 49 / (best_ms * 1e-3), "dtype": "f64", "gpu_launches_per_step": 1,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                     "kernel": "arz_step_kernel<2>", "kernel_ms": k_ms, "bytes_per_launch": N * bytes_per_env, "traffic": None,
                     "note": "1 substep per control step (SURVEY 8(d): HBM-relevant, 51 B per point update); at 16,384 envs the 40 MB "
                             "working set is L2-resident and the ~15 us launch is latency-bound, so the HBM fraction understates the kernel",
                     "peak_source": peaks["hbm_src"]},
        "cpu_baseline": _cpu("arz", cpu_seconds, k_envs=4),
        "e2e": {"value": N / (e2e_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": N * 8,
                "d2h_bytes_per_step": N * (102 * 8 + 8 + 2), "api": "env.step_host -> pdecg_arz_step_host (pinned host buffers)"},
        "clocks": sampler.window(t0, t1),
    }


def secondary_bt(torch, pcg, dev, sampler, peaks, cpu_seconds, N=16384):
    """Config 3b: BrainTumor1D x 16,384, notebook defaults (total_dosage 61.2), raw daily stepping, D / rho jittered +-10 %."""
    from oracle.cpu_arm import BT_P, bt_ic
    rng = np.random.default_rng(0)
    kw = dict(reward_class=pcg.BrainTumorReward(), D=0.2 * rng.uniform(0.9, 1.1, N), rho=0.03 * rng.uniform(0.9, 1.1, N), **BT_P)
    ic = {"init_condition": bt_ic()}
    env = pcg.make("PDEControlGym-BrainTumor1D", num_envs=N, device=dev, autoreset=True, **kw)
    env.reset(options=ic)
    acts = torch.empty(16, N, dtype=torch.float64, device=dev).uniform_(0, 0.1)
    eager_ms, _, _ = _event_loop(torch, dev, lambda i: env.step(acts[i % 16]), 300, 20)
    envg = pcg.make("PDEControlGym-BrainTumor1D", num_envs=N, device=dev, autoreset=True, **kw)
    envg.reset(options=ic)
    ms, _, t0, t1 = _graph_ms(torch, dev, pcg, envg, acts, lambda: envg.reset(options=ic))      # see secondary_arz
    envk = pcg.make("PDEControlGym-BrainTumor1D", num_envs=N, device=dev, copy_obs=False, **kw)
    envk.reset(options=ic)
    k_ms = _kernel_ms(torch, dev, lambda i: envk.step(acts[i % 16]), 200)
    envk.reset(options=ic)
    k_ms = min(k_ms, _graph_ms(torch, dev, pcg, envk, acts, lambda: envk.reset(options=ic), steps=200)[1])
    envh = pcg.make("PDEControlGym-BrainTumor1D", num_envs=N, device=dev, autoreset=True, **kw)
    envh.reset(options=ic)
    ah = _pinned(torch, 16, N, dtype=torch.float64).uniform_(0, 0.1)
    oh, rh = _pinned(torch, N, 201, dtype=torch.float64), _pinned(torch, N, dtype=torch.float64)
    teh, trh, vh = (_pinned(torch, N, dtype=torch.uint8) for _ in range(3))
    e2e_ms = _host_loop(torch, dev, lambda i: envh.step_host(ah[i % 16], oh, rh, teh, trh, vh), 100, 5)
    best_ms = min(ms, eager_ms)
    bytes_per_env = 8 * (201 + 201) + 2 * 64 + 2 * 32 + 8 + 8 + 3          # row in + out, ienv / fenv in + out, action, reward, flags
    gbs = N * bytes_per_env / (k_ms * 1e-3) / 1e9
    return {
        "config": "configs[2] (BrainTumor1D)", "workload": f"BrainTumor1D x{N} envs, 201 pts, one day per step, notebook defaults, D / rho jittered, f64",
        "metric": "env_steps_per_sec", "value": N / (best_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": best_ms, "n_envs": N,
        "api": "the faster of the two public stepping APIs on this box (autoreset=True, private observation copies): " + ("graph_replay" if ms <= eager_ms else "eager_step"),
        "eager_step": {"ms_per_step": eager_ms, "value": N / (eager_ms * 1e-3), "api": "env.step(action): one Python + ctypes dispatch per step (host-bound at this batch: moves with the box's CPU)"},
        "graph_replay": {"ms_per_step": ms, "value": N / (ms * 1e-3), "api": "pcg.GraphedStep(env, action).step(action): action copy + CUDA-graph replay of env.step"},
        "grid_pt_updates_per_sec": N * 199 / (best_ms * 1e-3), "dtype": "f64", "gpu_launches_per_step": 1,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                     "kernel": "bt_run_kernel<7, false>", "kernel_ms": k_ms, "bytes_per_launch": N * bytes_per_env, "traffic": None,
                     "note": "one explicit step per launch (SURVEY 8(d): HBM-relevant, 16 B per point update); at 16,384 envs the 54 MB "
                             "working set is L2-resident and the launch is latency-bound",
                     "peak_source": peaks["hbm_src"]},
        "cpu_baseline": _cpu("bt", cpu_seconds, k_envs=4),
        "e2e": {"value": N / (e2e_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": N * 8,
                "d2h_bytes_per_step": N * (201 * 8 + 8 + 3), "api": "env.step_host -> pdecg_bt_step_host (pinned host buffers)"},
        "clocks": sampler.window(t0, t1),
    }


def secondary_ns(torch, pcg, dev, sampler, peaks, cpu_seconds, N=1024):
    """Config 4: NavierStokes2D x 1,024, 21 x 21, 2000 Jacobi sweeps per step, NSReward."""
    kw = dict(boundary_condition=NS_BC, U_ref=np.zeros((200, 21, 21, 2)), action_ref=2.0 * np.ones(1000), reward_class=pcg.NSReward(0.1),
              action_dim=1, T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05)
    c = torch.empty(3, N, 1, 1, dtype=torch.float64, device=dev).uniform_(-5, 5) * torch.ones(1, 21, 21, dtype=torch.float64, device=dev)
    init = {"init": (c[0], c[1], c[2])}
    env = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=dev, autoreset=True, **kw)
    env.reset(options=init)
    acts = torch.empty(32, N, dtype=torch.float64, device=dev).uniform_(2, 4)
    ms, t0, t1 = _event_loop(torch, dev, lambda i: env.step(acts[i % 32]), 30, 3)
    envk = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=dev, copy_obs=False, **kw)
    envk.reset(options=init)
    k_ms = _kernel_ms(torch, dev, lambda i: envk.step(acts[i % 32]), 20)
    envh = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=dev, autoreset=True, **kw)
    envh.reset(options=init)
    ah = _pinned(torch, 32, N, dtype=torch.float64).uniform_(2, 4)
    oh, rh = _pinned(torch, N, 21, 21, 2, dtype=torch.float64), _pinned(torch, N, dtype=torch.float64)
    teh, trh, vh = (_pinned(torch, N, dtype=torch.uint8) for _ in range(3))
    e2e_ms = _host_loop(torch, dev, lambda i: envh.step_host(ah[i % 32], oh, rh, teh, trh, vh), 20, 3)
    units = N * 361 * 2000                                 # Jacobi point-iterations per launch
    tf6, tf5 = units * 6 / (k_ms * 1e-3) / 1e12, units * 5 / (k_ms * 1e-3) / 1e12
    return {
        "config": "configs[3]", "workload": f"NavierStokes2D x{N} envs, 21x21, 2000 Jacobi sweeps/step, NSReward, f64",
        "metric": "env_steps_per_sec", "value": N / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms, "n_envs": N,
        "grid_pt_updates_per_sec": units / (ms * 1e-3), "grid_pt_unit": "Jacobi point-iterations/s", "dtype": "f64",
        "gpu_launches_per_step": 1,
        "roofline": {"bound": "fp64", "achieved": tf6, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s", "frac": tf6 / peaks["fp64_tflops"],
                     "frac_issued": tf5 / peaks["fp64_tflops"], "kernel": "ns_step_warp_kernel (21x21)", "kernel_ms": k_ms, "traffic": None,
                     "flops_per_launch": units * 6,
                     "note": "6 flops per Jacobi point-iteration as the reference evaluates them (SURVEY 8(d)); the kernel issues 5 (b = dx*dy*rhs "
                             "hoisted): frac_issued.  DRAM traffic is ~0 (fields live in registers / shared memory for the whole step)",
                     "peak_source": "in-process non-FMA DADD/DMUL issue probe (pdecg_probe_fp_issue)"},
        "cpu_baseline": _cpu("ns", cpu_seconds, k_envs=1),
        "e2e": {"value": N / (e2e_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": N * 8,
                "d2h_bytes_per_step": N * (882 * 8 + 8 + 3), "api": "env.step_host -> pdecg_ns_step_host (pinned host buffers)"},
        "clocks": sampler.window(t0, t1),
    }


def secondary_ppo_rollout(torch, dist, pcg, dev, sampler, rank, world, env_kind, total_envs=1 << 20, rollout=16):
    """Config 5: PPO-style rollout collection (policy MLP 64-64 tanh -> action -> env.step, `rollout` steps) on 2^20 envs
    sharded over the N GPUs (strong scaling), one NCCL all-gather of episode statistics per rollout."""
    lo, hi = pcg.shard_range(total_envs, rank, world)
    n = hi - lo
    if env_kind == "rd":
        eid, npts, pts = "PDEControlGym-ReactionDiffusionPDE1D", 201, 199 * 100
        kw = dict(T=1.0, dt=1e-5, X=1.0, dx=5e-3, control_sample_rate=1e-3, reward_class=pcg.TunedReward1D(100000, -1e3, 3e2))
        beta = (50 * np.cos(8 * np.arccos(np.linspace(0, 1, npts)))).astype(np.float32)
    else:
        eid, npts, pts = "PDEControlGym-TransportPDE1D", 100, 99 * 1000
        kw = dict(T=5.0, dt=1e-4, X=1.0, dx=1e-2, control_sample_rate=0.1, reward_class=pcg.TunedReward1D(50000, -1e3, 3e2))
        beta = (5 * np.cos(7.35 * np.arccos(np.linspace(0, 1, npts)))).astype(np.float32)
    env = pcg.make(eid, num_envs=n, device=dev, precision="ref", normalize=True, max_control_value=20, limit_pde_state_size=True,
                   max_state_value=1e10, autoreset=True, copy_obs=False, env_id_offset=lo, **kw)
    obs, _ = env.reset(seed=0, options={"init_condition": pcg.UniformScaleIC(1, 10), "beta": beta})
    torch.manual_seed(0)                       # the same policy replica on every rank
    policy = torch.nn.Sequential(torch.nn.Linear(npts, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                                 torch.nn.Linear(64, 1)).to(dev)
    stats = pcg.EpisodeStats(n, dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1 + rank)
    box = {"obs": obs}

    @torch.no_grad()
    def one_rollout():
        o = box["obs"]
        for _ in range(rollout):
            mean = policy(o)
            act = (mean + torch.randn(mean.shape, device=dev, generator=gen)).clamp_(-1, 1)
            o, rew, term, trunc, _ = env.step(act)
            stats.update(rew, term, trunc)
        box["obs"] = o
        return stats.gather()                  # the only collective: [world, 9] doubles

    one_rollout()
    torch.cuda.synchronize(dev)
    times, t0 = [], time.time()
    for _ in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        one_rollout()
        e.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    t1 = time.time()
    ms = float(np.min(times))
    del env, policy
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    rate = total_envs * rollout / (ms * 1e-3)
    return {"config": "configs[4]", "workload": f"PPO rollout collection on {eid.split('-')[1]}: {total_envs} envs total over {world} GPU(s) "
                                                f"({n} on rank 0), {rollout} steps per rollout, 64-64 tanh policy, ref precision (f32), fused auto-reset with Philox ICs",
            "metric": "env_steps_per_sec", "value": rate, "unit": "env-steps/s", "ms_per_rollout": ms, "n_gpus": world, "scaling": "strong",
            "grid_pt_updates_per_sec": rate * pts, "dtype": "f32", "collectives_per_rollout": 1 if world > 1 else 0,
            "clocks": sampler.window(t0, t1) if sampler is not None else None}


# ---- our arm ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import pdecontrolgym_b200 as pcg
    from pdecontrolgym_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    E, K, W = args.envs, args.steps, args.warmup
    L = _lib.lib()

    def make_env(**extra):
        env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=E, device=dev, precision="fp64",
                       reward_class=pcg.NormReward(REWARD_NT, norm="2", horizon="temporal"), autoreset=True, **extra, **RD_KW)
        g = torch.Generator(device=dev)
        g.manual_seed(1234 + rank)          # rank-specific stream: global env id = rank*E + i
        c = torch.empty(E, 1, dtype=torch.float64, device=dev).uniform_(1, 10, generator=g)
        ic = c * torch.ones(1, NX_PTS, dtype=torch.float64, device=dev)
        env.reset(options={"init_condition": ic, "beta": beta_fn()})
        return env, g

    # ---------------- device-resident measurement (`value`) ----------------
    env, g = make_env()
    actions = torch.empty(K + W, E, dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g)
    ret_sum = torch.zeros(E, dtype=torch.float64, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    env._kernel_events = None
    n_envs_t = torch.tensor(float(E), device=dev, dtype=torch.float64)

    def rollout_stats():
        # once per rollout: all-gather of per-rank episode statistics (the only collective)
        stats = torch.stack([ret_sum.sum(), ret_sum.min(), ret_sum.max(), n_envs_t])
        if world > 1:
            gathered = torch.empty(world * stats.numel(), dtype=stats.dtype, device=dev)
            dist.all_gather_into_tensor(gathered, stats)
            return gathered
        return stats

    for i in range(W):                       # warm-up runs every op of the timed loop (lazy module loads)
        obs, rew, te, tr, _ = env.step(actions[i])
        ret_sum += rew
    rollout_stats()
    ret_sum.zero_()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    l0 = L.pdecg_launch_count()
    t_wall0 = time.time()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    t_issue0 = time.perf_counter()
    for i in range(K):
        env._kernel_events = ev[i]
        obs, rew, te, tr, _ = env.step(actions[W + i])
        ret_sum += rew
    env._kernel_events = None
    t_issue = time.perf_counter() - t_issue0
    rollout_stats()
    stop.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    launches = int(L.pdecg_launch_count() - l0)
    ms_total = start.elapsed_time(stop)
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    gaps = [ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(K - 1)]   # non-kernel time between step kernels
    if rank == 0 and gaps:
        gs = sorted(gaps)
        print(f"[bench] step-kernel ms: median {np.median(kern_ms):.4f} max {max(kern_ms):.4f}; inter-kernel gap ms: "
              f"median {gs[len(gs)//2]:.4f} p90 {gs[int(0.9*len(gs))]:.4f} max {gs[-1]:.4f} sum {sum(gs):.2f}; "
              f"host issue {1e3*t_issue/K:.4f} ms/step; head {start.elapsed_time(ev[0][0]):.3f} ms tail {ev[-1][1].elapsed_time(stop):.3f} ms; "
              f"big gaps {[(i, round(g, 2)) for i, g in enumerate(gaps) if g > 1.0]}", file=sys.stderr)
    clocks = sampler.window(t_wall0, t_wall1) if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    del env, obs
    torch.cuda.empty_cache()

    # ---------------- end-to-end measurement (`e2e`): host buffers through the C ABI ----------------
    # Observations cross PCIe as float32 -- the dtype of the reference's observation_space (base_env_1d.py:33-35) --
    # written by the step kernel itself (obs_dtype="float32"); the resident state stays float64.  The float64-observation
    # variant is timed next to it (`e2e.f64_obs`).
    def e2e_leg(obs_dtype):
        env2, _ = make_env(obs_dtype=obs_dtype)
        tdt = torch.float32 if obs_dtype == "float32" else torch.float64
        act_host = torch.empty(K + W, E, dtype=torch.float32).uniform_(-1, 1).pin_memory()
        obs_host = _pinned(torch, E, NX_PTS, dtype=tdt)
        rew_host = _pinned(torch, E, dtype=torch.float64)
        te_host, tr_host = _pinned(torch, E, dtype=torch.uint8), _pinned(torch, E, dtype=torch.uint8)
        for i in range(W):
            env2.step_host(act_host[i], obs_host, rew_host, te_host, tr_host)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        s2.record()
        acc = 0.0
        for i in range(K):
            env2.step_host(act_host[W + i], obs_host, rew_host, te_host, tr_host)
            acc += float(rew_host[0])            # the host consumes the result every step
        e2.record()
        torch.cuda.synchronize(dev)
        tw1 = time.perf_counter()
        ms = max(s2.elapsed_time(e2), 1e3 * (tw1 - tw0))   # host-side sync is part of the call
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        d2h = E * NX_PTS * obs_host.element_size() + E * 8 + 2 * E
        # D2H ceiling of this box at this N: every rank copies the same number of bytes device -> pinned host, back to
        # back, all ranks at once (what the copy-out of an end-to-end step could reach with a free kernel)
        src = torch.empty(E, NX_PTS, dtype=tdt, device=dev)
        reps = 20
        for _ in range(3):
            obs_host.copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(reps):
            obs_host.copy_(src, non_blocking=True)
        c1.record()
        torch.cuda.synchronize(dev)
        tc = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        ceiling = world * src.numel() * src.element_size() * reps / (float(tc.item()) * 1e-3) / 1e9
        del env2, src
        torch.cuda.empty_cache()
        return float(tt.item()), d2h, ceiling

    ms_e2e_max, d2h, d2h_ceiling = e2e_leg("float32")
    ms_e2e64_max, d2h64, d2h_ceiling64 = e2e_leg(None)
    h2d = E * 4

    # ---------------- config 5 at this N (every rank takes part) ----------------
    ppo = []
    if not args.no_secondary:
        for kind in ("rd", "transport"):
            try:
                r = secondary_ppo_rollout(torch, dist, pcg, dev, sampler if rank == 0 else None, rank, world, kind)
            except Exception as exc:  # noqa: BLE001
                r = {"config": "configs[4]", "workload": f"PPO rollout collection ({kind})", "error": repr(exc)[:300]} if rank == 0 else None
            if r is not None:
                ppo.append(r)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel ----------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    fp64_peak_tflops = _probe_peak(torch, dev, L, _lib, _lib.F64)
    sm_max_mhz = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    fp64_nominal = 148 * 64 * sm_max_mhz * 1e6 / 1e12           # 64 FP64 lanes per SM per clock, non-FMA
    k_ms = float(np.mean(kern_ms))
    flops_per_launch = E * INTERIOR * SUBSTEPS * FLOPS_PER_POINT
    # algorithmic HBM bytes per launch: state row in + out, the caller's private observation row out (obs_copy),
    # action, reward, 2 flags, time_index r/w, carry r/w (2 doubles)
    bytes_per_env = 3 * NX_PTS * 8 + 4 + 8 + 2 + 8 + 2 * 2 * 8
    bytes_per_launch = E * bytes_per_env + NX_PTS * 8
    traffic = None                # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture
    try:
        with open(os.path.join(ROOT, "profiles", "r2_rd_traffic.json")) as f:
            traffic = float(json.load(f)["dram_bytes_per_launch"]) if E == 65536 else None
    except Exception:  # noqa: BLE001
        pass
    achieved_tflops = flops_per_launch / (k_ms * 1e-3) / 1e12
    achieved_gbs = bytes_per_launch / (k_ms * 1e-3) / 1e9
    # issue-slot view: FP64 instructions the kernel issues (6 per point) on the lane slots it occupies (208 per 199 points)
    issue_frac = (E * 208 * SUBSTEPS * FP64_INSTR_PER_POINT / (k_ms * 1e-3) / 1e12) / fp64_peak_tflops
    useful_issue_frac = (E * INTERIOR * SUBSTEPS * FP64_INSTR_PER_POINT / (k_ms * 1e-3) / 1e12) / fp64_peak_tflops

    # ---------------- CPU baseline (bounded sample): the reference's own classes ----------------
    cores = len(os.sched_getaffinity(0))
    cpu_seconds = args.cpu_seconds
    if world == 1 and not args.no_cpu_baseline:
        cpu_baseline = _cpu("rd", 3 * cpu_seconds, k_envs=1)
        # the same integrator as compiled C with OpenMP on all host cores (oracle/c/pdecg_oracle.c): bounded sample
        try:
            from oracle.c_oracle import Bc1dCOracle
            n_c, steps_c = 16384, 40
            co = Bc1dCOracle("rd", precision="fp64", T=RD_KW["T"], dt=RD_KW["dt"], X=RD_KW["X"], dx=RD_KW["dx"],
                             control_sample_rate=RD_KW["control_sample_rate"], normalize=True, max_control_value=20)
            rng = np.random.default_rng(0)
            co.reset(rng.uniform(1, 10, (n_c, 1)) * np.ones((n_c, NX_PTS)), beta_fn())
            co.step(rng.uniform(-1, 1, n_c).astype(np.float32))
            t0 = time.perf_counter()
            for _ in range(steps_c):
                co.step(rng.uniform(-1, 1, n_c).astype(np.float32))
            dt_c = time.perf_counter() - t0
            cpu_baseline["c_port"] = {"value": n_c * steps_c / dt_c, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                      "sample": f"oracle/c/pdecg_oracle.c (gcc -O3 -mavx2 -fopenmp, no contraction; state integrator only), {n_c} envs x {steps_c} control steps, {dt_c:.2f} s wall"}
        except Exception as exc:  # noqa: BLE001
            cpu_baseline["c_port"] = {"unavailable": str(exc)[:200]}
    else:
        cpu_baseline = None

    # ---------------- secondary configurations (N = 1 only: they are single-GPU configurations) ----------------
    secondary = []
    if world == 1 and not args.no_secondary:
        pk = {"hbm_gbs": hbm_peak, "hbm_src": hbm_src, "fp64_tflops": fp64_peak_tflops, "sm_max_mhz": sm_max_mhz}
        for fn in (secondary_transport1, secondary_arz, secondary_bt, secondary_ns):
            try:
                secondary.append(fn(torch, pcg, dev, sampler, pk, cpu_seconds))
            except Exception as exc:  # noqa: BLE001
                secondary.append({"config": fn.__name__, "error": repr(exc)[:300]})
            torch.cuda.empty_cache()
    secondary.extend(ppo)
    sampler.stop()

    env_steps = world * E * K
    value = env_steps / (ms_total_max * 1e-3)
    e2e_value = world * E * K / (ms_e2e_max * 1e-3)
    d2h_achieved = world * d2h * K / (ms_e2e_max * 1e-3) / 1e9
    line = {
        "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s",
        "grid_pt_updates_per_sec": value * INTERIOR * SUBSTEPS,
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total_max / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": headline_config(E, world),
        "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                     "frac": achieved_tflops / fp64_peak_tflops, "traffic": traffic,
                     "traffic_source": "profiles/r2_rd_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full launch)",
                     "kernel": "bc1d_step_kernel<RD, ModeFp64, K=13, fast, 16 lanes/env>", "kernel_ms": k_ms,
                     "kernel_share_of_step": k_ms * K / ms_total,
                     "flops_per_launch": flops_per_launch,
                     "peak_source": "in-process non-FMA DADD/DMUL issue probe (pdecg_probe_fp_issue)",
                     "peak_nominal": fp64_nominal, "frac_of_nominal": achieved_tflops / fp64_nominal,
                     # the 7 flops/pt above count the exact FMA as 2 against a non-FMA peak; as issue slots:
                     "issue_slot_frac": issue_frac, "useful_issue_slot_frac": useful_issue_frac,
                     "issue_note": "kernel issues 6 FP64 instructions per point on 208 lane slots per 199 points (two envs per warp)",
                     "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                             "bytes_per_launch": bytes_per_launch, "peak_source": hbm_src}},
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e_max / K, "obs_dtype": "float32",
                "api": "env.step_host -> pdecg_bc1d_step_host (pinned host buffers, 4-stream chunk pipeline; float32 observations "
                       "written by the step kernel, float64 state resident)",
                "d2h_achieved_gbs": d2h_achieved, "d2h_ceiling_gbs": d2h_ceiling, "frac_of_d2h_ceiling": d2h_achieved / d2h_ceiling,
                "d2h_ceiling_how": f"all {world} rank(s) at once: 20 back-to-back pinned D2H copies of one step's observation bytes each, max over ranks",
                "f64_obs": {"value": world * E * K / (ms_e2e64_max * 1e-3), "ms_per_step": ms_e2e64_max / K, "d2h_bytes_per_step": d2h64,
                            "d2h_achieved_gbs": world * d2h64 * K / (ms_e2e64_max * 1e-3) / 1e9, "d2h_ceiling_gbs": d2h_ceiling64}},
        "gpu_launches": launches,
        "clocks": clocks,
        "secondary": secondary,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--cpu-seconds", type=float, default=3.0, help="CPU-baseline time budget per configuration")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

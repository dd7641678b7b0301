#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched PDE-env hot path (one JSON line on stdout).

Workload (BASELINE.json configs[1]): ReactionDiffusionPDE1D x 65,536 envs per GPU, reference default
grid (T=1, dt=1e-5, dx=5e-3 -> 201 points, 100 substeps per control step), Dirichlet boundary control,
NormReward, fresh random actions every step, fp64 state and arithmetic.  A "step" is ONE batched
control step of all envs (= 65,536 env-steps = 1.304e9 grid-point updates per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E]

N > 1: launched by torchrun, one rank per GPU, env batch sharded (weak scaling: E envs per GPU), no
collective inside the step; one NCCL all-gather of per-rank episode statistics per timed region.

`--impl reference`: the reference's own CPU path for the same workload -- the numpy oracle (a
restatement of the reference's numpy code; the reference itself is Python and cannot travel to the GPU
box) stepped in lockstep by an AsyncVectorEnv-equivalent (one worker process per host core).
"""
import argparse
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

# ---- the workload -------------------------------------------------------------------------------
NX_PTS = 201
SUBSTEPS = 100
INTERIOR = 199                       # points updated per substep (parabolic.py:143: u[1:Nx])
FLOPS_PER_POINT = 7                  # sub, mul(2*u), add, mul(F), add, mul(c*u), add -- dt*beta hoisted
RD_KW = dict(T=1.0, dt=1e-5, X=1.0, dx=5e-3, control_sample_rate=1e-3, normalize=True,
             max_control_value=20, sensing_loc="full", control_type="Dirchilet", sensing_type=None,
             limit_pde_state_size=False)
REWARD_NT = 100000


def workload_name(envs):
    return (f"ReactionDiffusionPDE1D x{envs} envs/GPU, 201 pts, 100 substeps/step, Dirichlet, NormReward(L2,temporal), "
            "random actions, fp64")


def beta_fn():
    return 50 * np.cos(8 * np.arccos(np.linspace(0, 1, NX_PTS)))


# ---- clocks sampler ------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (NVML in a thread; the
    `nvidia-smi -lms` loop of the profiling recipe is only the fallback because each of its queries
    stalls kernel launches for milliseconds and distorts a sub-second timed region)."""
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period=0.05):
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

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.proc is not None:
            time.sleep(0.6)
            self.proc.terminate()
        if self.thread is not None:
            self.thread.join(timeout=1.0)
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


# ---- CPU baseline: AsyncVectorEnv-equivalent over the numpy oracle -----------------------------------
def _cpu_worker(conn, k_envs, seed):
    from oracle import np_oracle as O
    rng = np.random.default_rng(seed)
    envs = []
    for _ in range(k_envs):
        e = O.ReactionDiffusionOracle(1, reward=O.NormRewardSpec(REWARD_NT, norm=2), precision="fp64",
                                      **{k: v for k, v in RD_KW.items()})
        e.reset(np.ones(NX_PTS) * rng.uniform(1, 10), beta_fn())
        envs.append(e)
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg is None:
            break
        out = []
        for e, a in zip(envs, msg):
            o, r, te, tr = e.step(np.array([a], dtype=np.float32))
            if te[0] or tr[0]:         # same-step auto-reset like the vector env (a finished env would step for free)
                o = e.reset(np.ones(NX_PTS) * rng.uniform(1, 10), beta_fn())
            out.append((o, float(r[0]), bool(te[0]), bool(tr[0])))
        conn.send(out)


def cpu_lockstep_rate(steps, warmup, cores=None, k_envs=1):
    """env-steps/s of `cores` worker processes x k_envs oracle envs stepped in lockstep through pipes."""
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    conns, procs = [], []
    for w in range(cores):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_cpu_worker, args=(b, k_envs, 1000 + w), daemon=True)
        p.start()
        conns.append(a)
        procs.append(p)
    for c in conns:
        assert c.recv() == "ready"
    rng = np.random.default_rng(0)

    def one_step():
        acts = rng.uniform(-1, 1, (cores, k_envs)).astype(np.float32)
        for c, a in zip(conns, acts):
            c.send(list(a))
        for c in conns:
            c.recv()

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    for c in conns:
        c.send(None)
    for p in procs:
        p.join(timeout=5)
    n_env_steps = steps * cores * k_envs
    return n_env_steps / dt, dt, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    # each bench "step" = one lockstep control step of (cores x 4) reference-style envs
    k_envs = 4
    sub_steps_per_step = 25           # bounded sample: 25 lockstep rounds per reported step
    rate, dt, cores = cpu_lockstep_rate(args.steps * sub_steps_per_step, args.warmup * sub_steps_per_step, cores, k_envs)
    sample = (f"numpy oracle (port of the reference numpy path), {cores} worker processes x {k_envs} envs in lockstep over pipes "
              f"(AsyncVectorEnv-equivalent), {args.steps * sub_steps_per_step} control steps each")
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": rate, "unit": "env-steps/s",
        "grid_pt_updates_per_sec": rate * INTERIOR * SUBSTEPS,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.envs), "cpu_envs": cores * k_envs},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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

    def make_env():
        env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=E, device=dev, precision="fp64",
                       reward_class=pcg.NormReward(REWARD_NT, norm="2", horizon="temporal"), autoreset=True, **RD_KW)
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
    def rollout_stats():
        # once per rollout: all-gather of per-rank episode statistics (the only collective)
        stats = torch.stack([ret_sum.sum(), ret_sum.min(), ret_sum.max(), n_envs_t])
        if world > 1:
            gathered = torch.empty(world * stats.numel(), dtype=stats.dtype, device=dev)
            dist.all_gather_into_tensor(gathered, stats)
            return gathered
        return stats

    n_envs_t = torch.tensor(float(E), device=dev, dtype=torch.float64)
    for i in range(W):                       # warm-up runs every op of the timed loop (lazy module loads)
        obs, rew, te, tr, _ = env.step(actions[i])
        ret_sum += rew
    rollout_stats()
    ret_sum.zero_()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
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
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())

    # ---------------- end-to-end measurement (`e2e`): host buffers through the C ABI ----------------
    env2, g2 = make_env()
    act_host = torch.empty(K + W, E, dtype=torch.float32).uniform_(-1, 1).pin_memory()
    obs_host = torch.empty(E, NX_PTS, dtype=torch.float64).pin_memory()
    rew_host = torch.empty(E, dtype=torch.float64).pin_memory()
    te_host = torch.empty(E, dtype=torch.uint8).pin_memory()
    tr_host = torch.empty(E, dtype=torch.uint8).pin_memory()
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
    ms_e2e = max(s2.elapsed_time(e2), 1e3 * (tw1 - tw0))   # host-side sync is part of the call
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e_max = float(t.item())
    h2d = E * 4
    d2h = E * NX_PTS * 8 + E * 8 + 2 * E

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
    import ctypes as C
    ms = C.c_float()
    ops = C.c_double()
    _lib.check(L.pdecg_probe_fp_issue(_lib.F64, 148 * 8, 256, 20000, C.byref(ms), C.byref(ops), torch.cuda.current_stream(dev).cuda_stream))
    fp64_peak_tflops = ops.value / (ms.value * 1e-3) / 1e12
    k_ms = float(np.mean(kern_ms))
    flops_per_launch = E * INTERIOR * SUBSTEPS * FLOPS_PER_POINT
    # algorithmic HBM bytes per launch: state row in + out, the caller's private observation row out (obs_copy),
    # action, reward, 2 flags, time_index r/w, carry r/w (2 doubles)
    bytes_per_env = 3 * NX_PTS * 8 + 4 + 8 + 2 + 8 + 2 * 2 * 8
    bytes_per_launch = E * bytes_per_env + NX_PTS * 8
    traffic = None                # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture
    try:
        with open(os.path.join(ROOT, "profiles", "r1_rd_traffic.json")) as f:
            traffic = float(json.load(f)["dram_bytes_per_launch"]) if E == 65536 else None
    except Exception:  # noqa: BLE001
        pass
    achieved_tflops = flops_per_launch / (k_ms * 1e-3) / 1e12
    achieved_gbs = bytes_per_launch / (k_ms * 1e-3) / 1e9

    # ---------------- CPU baseline (bounded sample) ----------------
    cores = len(os.sched_getaffinity(0))
    if world == 1 and not args.no_cpu_baseline:
        cpu_rate, cpu_dt, cores = cpu_lockstep_rate(steps=400, warmup=10, cores=cores, k_envs=4)
        cpu_baseline = {"value": cpu_rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
                        "sample": f"numpy oracle (port of the reference numpy path): {cores} worker processes x 4 envs, 400 lockstep "
                                  f"control steps over pipes (AsyncVectorEnv-equivalent), {cpu_dt:.1f} s wall"}
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

    env_steps = world * E * K
    value = env_steps / (ms_total_max * 1e-3)
    line = {
        "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s",
        "grid_pt_updates_per_sec": value * INTERIOR * SUBSTEPS,
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total_max / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(E), "envs_per_gpu": E, "global_envs": world * E,
                   "parallelism": f"env-shard x{world} (no collective in step; 1 stats all-gather per timed region)",
                   "l2": "no flush: state (105 MB) + observation copy (105 MB) written every step exceed the 126 MB L2",
                   "autoreset": True},
        "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                     "frac": achieved_tflops / fp64_peak_tflops, "traffic": traffic,
                     "traffic_source": "profiles/r1_rd_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full launch)",
                     "kernel": "bc1d_step_kernel<RD, ModeFp64, K=13, fast, 16 lanes/env>", "kernel_ms": k_ms,
                     "kernel_share_of_step": k_ms * K / ms_total,
                     "flops_per_launch": flops_per_launch,
                     "peak_source": "in-process non-FMA DADD/DMUL issue probe (pdecg_probe_fp_issue)",
                     "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                             "bytes_per_launch": bytes_per_launch, "peak_source": hbm_src}},
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": world * E * K / (ms_e2e_max * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e_max / K,
                "api": "env.step_host -> pdecg_bc1d_step_host (pinned host buffers)"},
        "gpu_launches": launches,
        "clocks": clocks,
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
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

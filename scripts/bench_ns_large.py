"""NavierStokes2D on grids beyond the shared-memory-resident limit (the HBM-resident path of navier_stokes_large.cu):
device-resident step time and the roofline of its dominant kernel, the TMA-tiled temporally blocked Jacobi sweep.

    python scripts/bench_ns_large.py [--grids 128,256,512] [--iters 2000] [--kt 10]

Per grid one JSON line: env-steps/s, Jacobi point-iterations/s and `roofline` = {bound: "hbm", achieved, peak, frac, ...}
where achieved = ALGORITHMIC bytes per tiled launch (each interior point's p and b read once, p written once: 24 B per
point per launch of KT sweeps) / that launch's average duration.  The duration of one tiled launch is
(ms(step with `iters` sweeps) - ms(step with 1 sweep)) / number of tiled launches, all timed with CUDA events."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pdecontrolgym_b200 as pcg

BC = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"], "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}


def step_ms(n, N, iters, dev, reps):
    h = 1.0 / (n - 1)
    dt = 1e-3 * (h / 0.05) ** 2
    env = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=dev, boundary_condition=BC, U_ref=np.zeros((8, n, n, 2)), action_ref=2.0 * np.ones(1000),
                   reward_class=pcg.NSReward(0.1), action_dim=1, maximum_pressure_iteration=iters, T=1000 * dt, dt=dt, X=1, dx=h, Y=1, dy=h,
                   copy_obs=False)
    c = torch.empty(3, N, 1, 1, dtype=torch.float64, device=dev).uniform_(-5, 5) * torch.ones(1, n, n, dtype=torch.float64, device=dev)
    acts = torch.empty(4, N, dtype=torch.float64, device=dev).uniform_(2, 4)
    best = 1e30
    for rep in range(reps + 1):
        env.reset(options={"init": (c[0], c[1], c[2])})
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for k in range(3):
            env.step(acts[k])
        e.record()
        torch.cuda.synchronize()
        if rep > 0:
            best = min(best, s.elapsed_time(e) / 3)
    del env
    torch.cuda.empty_cache()
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grids", default="128,256,512")
    ap.add_argument("--iters", type=int, default=2000)
    ap.add_argument("--kt", type=int, default=10)
    ap.add_argument("--points", type=float, default=2.7e7, help="grid points per batch (sets the env count per grid)")
    args = ap.parse_args()
    os.environ["PDECG_NS_KT"] = str(args.kt)
    dev = torch.device("cuda:0")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    for n in [int(x) for x in args.grids.split(",")]:
        N = max(1, int(args.points // (n * n)))
        ms_full = step_ms(n, N, args.iters, dev, 2)
        ms_one = step_ms(n, N, 1, dev, 2)
        launches = -(-(args.iters - 1) // args.kt)
        k_ms = (ms_full - ms_one) / launches
        interior = (n - 2) ** 2
        bytes_launch = N * interior * 24
        gbs = bytes_launch / (k_ms * 1e-3) / 1e9
        units = N * interior * args.iters
        print(json.dumps({
            "config": f"NavierStokes2D {n}x{n}, {args.iters} Jacobi sweeps x{N} envs (HBM-resident path, {args.kt} sweeps per TMA round trip)",
            "n_envs": N, "ms_per_batched_step": ms_full, "ms_step_with_one_sweep": ms_one, "env_steps_per_sec": N / (ms_full * 1e-3),
            "grid_pt_updates_per_sec": units / (ms_full * 1e-3), "grid_pt_unit": "Jacobi point-iterations/s",
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "kernel": f"ns_lg_jacobi_tiled_kernel<{args.kt}>",
                         "kernel_ms": k_ms, "tiled_launches_per_step": launches, "bytes_per_launch": bytes_launch, "traffic": None,
                         "effective_gbs_per_sweep": N * interior * 24 * min(args.kt, args.iters - 1) / (k_ms * 1e-3) / 1e9,
                         "note": "algorithmic bytes per launch: p and b in, p out, once per interior point (24 B); `effective_gbs_per_sweep` is what an "
                                 "un-blocked sweep-per-launch kernel would have to move for the same work"}}), flush=True)


if __name__ == "__main__":
    main()

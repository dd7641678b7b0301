import os, sys, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import pdecontrolgym_b200 as pcg
dev = torch.device("cuda:0")
def timed(fn, steps, warmup=5):
    for i in range(warmup): fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps): fn(i)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / steps
for eid, npts, kw, prec, rew, bdt in (
    ("PDEControlGym-TransportPDE1D", 100, dict(T=1.0, dt=1e-4, X=1.0, dx=1e-2, control_sample_rate=0.1), "ref", pcg.TunedReward1D(10000, -1e3, 3e2), np.float32),
    ("PDEControlGym-ReactionDiffusionPDE1D", 201, dict(T=1.0, dt=1e-5, X=1.0, dx=5e-3, control_sample_rate=1e-3), "fp64", pcg.NormReward(100000), np.float64)):
    for N in (1, 256, 1024, 2048, 4096, 8192, 16384):
        out = {}
        for mode, thr in (("full", "1000000000"), ("half", "0")):
            os.environ["PDECG_BC1D_HALFWARP_MIN"] = thr
            env = pcg.make(eid, num_envs=N, device=dev, precision=prec, reward_class=rew, normalize=True, max_control_value=20, autoreset=True, copy_obs=False, **kw)
            ic = torch.empty(N, 1, dtype=torch.float64, device=dev).uniform_(1, 10) * torch.ones(1, npts, dtype=torch.float64, device=dev)
            env.reset(options={"init_condition": ic, "beta": (5 * np.cos(7.35 * np.arccos(np.linspace(0, 1, npts)))).astype(bdt)})
            acts = torch.empty(40, N, device=dev).uniform_(-1, 1)
            out[mode] = timed(lambda i: env.step(acts[i]), 30)
        print(json.dumps({"env": eid.split("-")[1], "N": N, "ms_full_warp": round(out["full"], 4), "ms_half_warp": round(out["half"], 4)}), flush=True)

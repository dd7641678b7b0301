import os, sys, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import pdecontrolgym_b200 as pcg
dev = torch.device("cuda:0")
bc = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"], "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}
for N in (64, 256, 1024, 2048, 4096, 8192):
    res = {}
    for mode in ("wpb1", "wpb4"):
        os.environ.pop("PDECG_NS_WPB4", None)
        os.environ.pop("PDECG_NS_WPB1", None)
        os.environ["PDECG_NS_WPB4" if mode == "wpb4" else "PDECG_NS_WPB1"] = "1"
        env = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=dev, boundary_condition=bc, U_ref=np.zeros((200, 21, 21, 2)), action_ref=2.0 * np.ones(1000),
                       reward_class=pcg.NSReward(0.1), action_dim=1, T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05, copy_obs=False)
        c = torch.empty(3, N, 1, 1, dtype=torch.float64, device=dev).uniform_(-5, 5) * torch.ones(1, 21, 21, dtype=torch.float64, device=dev)
        env.reset(options={"init": (c[0], c[1], c[2])})
        acts = torch.empty(32, N, dtype=torch.float64, device=dev).uniform_(2, 4)
        for i in range(3): env.step(acts[i])
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(20): env.step(acts[3 + i])
        e.record(); torch.cuda.synchronize()
        res[mode] = s.elapsed_time(e) / 20
    print(json.dumps({"N": N, **{k: round(v, 4) for k, v in res.items()}}), flush=True)

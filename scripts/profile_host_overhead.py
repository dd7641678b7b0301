"""cProfile of the host side of env.step() at a tiny batch (GPU work negligible)."""
import cProfile, pstats, os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pdecontrolgym_b200 as pcg
dev = torch.device("cuda:0")
env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=64, device=dev, precision="fp64", reward_class=pcg.NormReward(100000), autoreset=True,
               T=1.0, dt=1e-5, X=1.0, dx=5e-3, control_sample_rate=1e-3, normalize=True)
env.reset(options={"init_condition": np.ones(201), "beta": np.zeros(201)})
a = torch.zeros(64, device=dev)
for _ in range(50): env.step(a)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(2000): env.step(a)
torch.cuda.synchronize()
print("us per step (wall):", 1e6 * (time.perf_counter() - t0) / 2000)
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): env.step(a)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])

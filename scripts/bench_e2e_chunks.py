"""A/B of the host-buffer step's chunking (PDECG_HOST_CHUNKS): e2e ms per 65,536-env fp64 ReactionDiffusion step with
float32 observations.  Run once per setting (the cap is read once per process):  PDECG_HOST_CHUNKS=4 python scripts/bench_e2e_chunks.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pdecontrolgym_b200 as pcg

dev = torch.device("cuda:0")
E = 65536
env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=E, device=dev, precision="fp64", reward_class=pcg.NormReward(100000, norm="2"),
               autoreset=True, obs_dtype="float32", T=1.0, dt=1e-5, X=1.0, dx=5e-3, control_sample_rate=1e-3, normalize=True, max_control_value=20)
ic = torch.empty(E, 1, dtype=torch.float64, device=dev).uniform_(1, 10) * torch.ones(1, 201, dtype=torch.float64, device=dev)
env.reset(options={"init_condition": ic, "beta": 50 * np.cos(8 * np.arccos(np.linspace(0, 1, 201)))})
a = torch.empty(64, E, dtype=torch.float32).uniform_(-1, 1).pin_memory()
oh = torch.empty(E, 201, dtype=torch.float32).pin_memory()
rh = torch.empty(E, dtype=torch.float64).pin_memory()
th, uh = torch.empty(E, dtype=torch.uint8).pin_memory(), torch.empty(E, dtype=torch.uint8).pin_memory()
for i in range(10):
    env.step_host(a[i], oh, rh, th, uh)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50):
    env.step_host(a[i % 64], oh, rh, th, uh)
torch.cuda.synchronize()
print(f"PDECG_HOST_CHUNKS={os.environ.get('PDECG_HOST_CHUNKS', '8 (default)')}: {(time.perf_counter() - t0) / 50 * 1e3:.4f} ms per e2e step")

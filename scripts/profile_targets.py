"""Tiny drivers for ncu captures of the non-headline kernels:  python scripts/profile_targets.py arz|bt|bt_wrap|ns|transport|rd_large|tr_large|ns_large"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pdecontrolgym_b200 as pcg

dev = torch.device("cuda:0")
which = sys.argv[1]
if which == "arz":
    N = 262144
    env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=dev, simulation_type="outlet", reward_class=pcg.TrafficARZReward(),
                   T=240, dt=0.25, X=500, dx=10, copy_obs=False)
    env.reset()
    a = torch.empty(N, 1, dtype=torch.float64, device=dev).uniform_(0.96, 1.44)
    for _ in range(6):
        env.step(a)
elif which in ("bt", "bt_wrap"):
    N = 262144 if which == "bt" else 65536
    xs = np.linspace(0, 200, 201)
    env = pcg.make("PDEControlGym-BrainTumor1D", num_envs=N, device=dev, reward_class=pcg.BrainTumorReward(), total_dosage=61.2,
                   T=600, dt=1, X=200, dx=1, copy_obs=False)
    env.reset(options={"init_condition": 0.8 * 1e5 * np.exp(-0.25 * xs ** 2)})
    if which == "bt":
        a = torch.empty(N, dtype=torch.float64, device=dev).uniform_(0, 0.1)
        for _ in range(6):
            env.step(a)
    else:
        w = pcg.TherapyWrapper(env)
        w.benchmark()
        w.reset()
elif which == "ns":
    N = 1024
    bc = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"], "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}
    env = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=dev, boundary_condition=bc, U_ref=np.zeros((200, 21, 21, 2)),
                   action_ref=2.0 * np.ones(1000), reward_class=pcg.NSReward(0.1), action_dim=1, T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05, copy_obs=False)
    z = torch.zeros(N, 21, 21, dtype=torch.float64, device=dev)
    env.reset(options={"init": (z + 1.0, z - 2.0, z + 0.5)})
    a = torch.empty(N, dtype=torch.float64, device=dev).uniform_(2, 4)
    for _ in range(4):
        env.step(a)
elif which == "transport":
    N = 65536
    env = pcg.make("PDEControlGym-TransportPDE1D", num_envs=N, device=dev, precision="ref", reward_class=pcg.TunedReward1D(10000, -1e3, 3e2),
                   T=1.0, dt=1e-4, X=1.0, dx=1e-2, control_sample_rate=0.1, normalize=True, max_control_value=20, copy_obs=False)
    ic = torch.empty(N, 1, dtype=torch.float64, device=dev).uniform_(1, 10) * torch.ones(1, 100, dtype=torch.float64, device=dev)
    env.reset(options={"init_condition": ic, "beta": (5 * np.cos(7.35 * np.arccos(np.linspace(0, 1, 100)))).astype(np.float32)})
    a = torch.empty(N, device=dev).uniform_(-1, 1)
    for _ in range(4):
        env.step(a)
elif which in ("rd_large", "tr_large"):
    # grids beyond 512 points: the CTA-per-env shared-memory kernel (bc1d_large.cuh)
    N, nx = 8192, 1000
    dx = 1.0 / nx
    if which == "rd_large":
        kw = dict(T=1.0, dt=0.4 * dx * dx, X=1.0, dx=dx, control_sample_rate=100 * 0.4 * dx * dx)
        npts, eid, prec = nx + 1, "PDEControlGym-ReactionDiffusionPDE1D", "fp64"
    else:
        kw = dict(T=1.0, dt=dx / 100, X=1.0, dx=dx, control_sample_rate=10 * dx)
        npts, eid, prec = nx, "PDEControlGym-TransportPDE1D", "ref"
    env = pcg.make(eid, num_envs=N, device=dev, precision=prec, reward_class=pcg.NormReward(int(round(kw["T"] / kw["dt"]))), normalize=True,
                   max_control_value=20, copy_obs=False, **kw)
    ic = torch.empty(N, 1, dtype=torch.float64, device=dev).uniform_(1, 10) * torch.ones(1, npts, dtype=torch.float64, device=dev)
    beta = 5 * np.cos(7.35 * np.arccos(np.linspace(0, 1, npts)))
    env.reset(options={"init_condition": ic, "beta": beta.astype(np.float64 if prec == "fp64" else np.float32)})
    a = torch.empty(N, device=dev).uniform_(-1, 1)
    for _ in range(4):
        env.step(a)
elif which == "ns_large":
    N, n = 592, 41
    bc = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"], "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}
    env = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=dev, boundary_condition=bc, U_ref=np.zeros((200, n, n, 2)),
                   action_ref=2.0 * np.ones(1000), reward_class=pcg.NSReward(0.1), action_dim=1, T=0.05, dt=2.5e-4, X=1, dx=0.025, Y=1, dy=0.025, copy_obs=False)
    z = torch.zeros(N, n, n, dtype=torch.float64, device=dev)
    env.reset(options={"init": (z + 1.0, z - 2.0, z + 0.5)})
    a = torch.empty(N, dtype=torch.float64, device=dev).uniform_(2, 4)
    for _ in range(3):
        env.step(a)
elif which == "ns_tiled":
    # 512 x 512: the HBM-resident path (navier_stokes_large.cu); 101 sweeps = 1 plain + 10 tiled launches of 10
    N, n = 96, 512
    h = 1.0 / (n - 1)
    dt = 1e-3 * (h / 0.05) ** 2
    bc = {"upper": ["Controllable", "Dirchilet"], "lower": ["Dirchilet", "Dirchilet"], "left": ["Dirchilet", "Dirchilet"], "right": ["Dirchilet", "Dirchilet"]}
    env = pcg.make("PDEControlGym-NavierStokes2D", num_envs=N, device=dev, boundary_condition=bc, U_ref=np.zeros((8, n, n, 2)),
                   action_ref=2.0 * np.ones(1000), reward_class=pcg.NSReward(0.1), action_dim=1, maximum_pressure_iteration=101, T=100 * dt, dt=dt,
                   X=1, dx=h, Y=1, dy=h, copy_obs=False)
    c = torch.empty(3, N, 1, 1, dtype=torch.float64, device=dev).uniform_(-5, 5) * torch.ones(1, n, n, dtype=torch.float64, device=dev)
    env.reset(options={"init": (c[0], c[1], c[2])})
    a = torch.empty(N, dtype=torch.float64, device=dev).uniform_(2, 4)
    for _ in range(2):
        env.step(a)
torch.cuda.synchronize()

"""Small-batch pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pdecontrolgym_b200 as pcg

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
for eid, npts, kw, rew in (
        ("PDEControlGym-TransportPDE1D", 100, dict(T=0.3, dt=1e-4, X=1, dx=1e-2, control_sample_rate=0.05), pcg.TunedReward1D(3000)),
        ("PDEControlGym-ReactionDiffusionPDE1D", 201, dict(T=0.003, dt=1e-5, X=1, dx=5e-3, control_sample_rate=1e-3), pcg.NormReward(300, "inf", "t-horizon"))):
    for prec in ("ref", "fp64", "fp32"):
        for ct, sl in (("Dirchilet", "full"), ("Neumann", "collocated")):
            env = pcg.make(eid, num_envs=7, device=dev, precision=prec, reward_class=rew, control_type=ct, sensing_loc=sl, autoreset=True, **kw)
            env.reset(options={"init_condition": rng.uniform(1, 2, (7, npts)), "beta": rng.uniform(-1, 1, npts)})
            for _ in range(4):
                env.step(rng.uniform(-1, 1, 7).astype(np.float32))
env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=9001, device=dev, precision="fp64", reward_class=pcg.NormReward(300),
               T=0.003, dt=1e-5, X=1, dx=5e-3, control_sample_rate=1e-3)
env.reset(options={"init_condition": np.ones(201), "beta": np.zeros(201)})
a = torch.zeros(9001).pin_memory()
o = torch.empty(9001, 201, dtype=torch.float64).pin_memory()
env.step_host(a, o, torch.empty(9001, dtype=torch.float64).pin_memory(), torch.empty(9001, dtype=torch.uint8).pin_memory(), torch.empty(9001, dtype=torch.uint8).pin_memory())
env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=5, device=dev, simulation_type="both", control_freq=2, T=5, dt=0.25, X=500, dx=10)
env.reset()
for _ in range(25):
    env.step(rng.uniform(1.0, 1.4, (5, 2)))
xs = np.linspace(0, 200, 201)
bt = pcg.make("PDEControlGym-BrainTumor1D", num_envs=5, device=dev, total_dosage=61.2, T=600, dt=1, X=200, dx=1)
bt.reset(options={"init_condition": 0.8e5 * np.exp(-0.25 * xs ** 2)})
w = pcg.TherapyWrapper(bt, weekends=True)
w.benchmark()
w.reset()
for _ in range(40):
    w.step(rng.uniform(0, 0.1, 5))
bc = {"upper": ["Controllable", "Neumann"], "lower": ["Neumann", "Dirchilet"], "left": ["Dirchilet", "Controllable"], "right": ["Neumann", "Neumann"]}
for n in (21, 12):
    h = 1.0 / (n - 1)
    ns = pcg.make("PDEControlGym-NavierStokes2D", num_envs=5, device=dev, boundary_condition=bc, U_ref=np.zeros((50, n, n, 2)), action_ref=np.ones(100),
                  reward_class=pcg.NSReward(), action_dim=1, maximum_pressure_iteration=20, T=0.02, dt=1e-3 * (h / 0.05) ** 2, X=1, dx=h, Y=1, dy=h)
    ns.reset(options={"init": tuple(rng.uniform(-1, 1, (5, n, n)) for _ in range(3))})
    for _ in range(3):
        ns.step(rng.uniform(2, 4, 5))
    ns.solve_pressure(rng.uniform(-1, 1, (5, n, n)), rng.uniform(-1, 1, (5, n, n)), np.zeros((n, n)))
# grids beyond 512 points: register-tiled and shared-memory CTA kernels, every reward event path, fused auto-reset
for eid, nx, sub in (("PDEControlGym-TransportPDE1D", 700, 30), ("PDEControlGym-ReactionDiffusionPDE1D", 2500, 30)):
    dx = 1.0 / nx
    dt = dx / 100 if "Transport" in eid else 0.4 * dx * dx
    npts = nx if "Transport" in eid else nx + 1
    for prec, rew, ct in (("ref", pcg.TunedReward1D(4 * sub), "Dirchilet"), ("fp64", pcg.NormReward(4 * sub, "1", "differential"), "Dirchilet"),
                          ("fp64", pcg.NormReward(4 * sub, "inf", "t-horizon"), "Neumann"), ("fp32", pcg.NormReward(4 * sub), "Dirchilet")):
        env = pcg.make(eid, num_envs=3, device=dev, precision=prec, reward_class=rew, control_type=ct, autoreset=True,
                       T=4 * sub * dt, dt=dt, X=1, dx=dx, control_sample_rate=sub * dt)
        env.reset(options={"init_condition": rng.uniform(1, 2, (3, npts)), "beta": rng.uniform(-1, 1, npts)})
        for _ in range(5):
            env.step(rng.uniform(-1, 1, 3).astype(np.float32))
os.environ["PDECG_BC1D_LARGE_SMEM"] = "1"
env = pcg.make("PDEControlGym-TransportPDE1D", num_envs=3, device=dev, reward_class=pcg.TunedReward1D(600), autoreset=True,
               T=600 * 1e-5, dt=1e-5, X=1, dx=1e-3, control_sample_rate=150 * 1e-5)
env.reset(options={"init_condition": rng.uniform(1, 2, (3, 1000)), "beta": rng.uniform(-1, 1, 1000)})
for _ in range(5):
    env.step(rng.uniform(-1, 1, 3).astype(np.float32))
del os.environ["PDECG_BC1D_LARGE_SMEM"]
# fused auto-reset of the other families, larger ARZ / BrainTumor / NavierStokes grids
env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=5, device=dev, simulation_type="outlet", T=1.5, dt=0.025, X=500, dx=1, autoreset=True)
env.reset()
for _ in range(70):
    env.step(rng.uniform(1.0, 1.4, (5, 1)))
xs = np.linspace(0, 300, 301)
bt = pcg.make("PDEControlGym-BrainTumor1D", num_envs=5, device=dev, total_dosage=61.2, T=20, dt=1, X=300, dx=1, autoreset=True)
bt.reset(options={"init_condition": 0.8e5 * np.exp(-0.25 * xs ** 2)})
for _ in range(45):
    bt.step(rng.uniform(0, 0.1, 5))
for n in (21, 33, 50):
    h = 1.0 / (n - 1)
    ns = pcg.make("PDEControlGym-NavierStokes2D", num_envs=3, device=dev, boundary_condition=bc, U_ref=np.zeros((50, n, n, 2)), action_ref=np.ones(100),
                  reward_class=pcg.NSReward(), action_dim=1, maximum_pressure_iteration=10, T=0.004 * (h / 0.05) ** 2, dt=1e-3 * (h / 0.05) ** 2, X=1, dx=h, Y=1, dy=h,
                  autoreset=True)
    ns.reset(options={"init": tuple(rng.uniform(-1, 1, (3, n, n)) for _ in range(3))})
    for _ in range(7):
        ns.step(rng.uniform(2, 4, 3))
# ---- round 2: float32-observation / final_obs paths, float32 modes, the HBM-resident NavierStokes path (TMA tiles),
#      the register-blocked and shared-memory CTA Jacobi, both ARZ lane mappings + its exact fallback, Neuron ----
env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=4099, device=dev, precision="fp64", reward_class=pcg.NormReward(3),
               T=0.003, dt=1e-5, X=1, dx=5e-3, control_sample_rate=1e-3, autoreset=True, obs_dtype="float32", final_obs=True)
env.reset(options={"init_condition": np.ones(201), "beta": np.zeros(201)})
ah = torch.zeros(4099).pin_memory()
oh = torch.empty(4099, 201, dtype=torch.float32).pin_memory()
for _ in range(4):
    env.step(torch.zeros(4099, device=dev))
    env.step_host(ah, oh, torch.empty(4099, dtype=torch.float64).pin_memory(), torch.empty(4099, dtype=torch.uint8).pin_memory(), torch.empty(4099, dtype=torch.uint8).pin_memory())
for lpe16 in (False, True):
    if lpe16:
        os.environ["PDECG_ARZ_LPE16"] = "1"
    for prec in ("ref", "fp32"):
        env = pcg.make("PDEControlGym-TrafficPDE1D", num_envs=5, device=dev, simulation_type="outlet", T=2, dt=0.25, X=500, dx=10, autoreset=True,
                       final_obs=True, precision=prec)
        env.reset()
        env.y[1, 7] = 0.0                                   # an exceptional operand: that warp redoes its step with IEEE divisions
        for _ in range(40):
            env.step(rng.uniform(1.0, 1.4, (5, 1)))
os.environ.pop("PDECG_ARZ_LPE16", None)
bt = pcg.make("PDEControlGym-BrainTumor1D", num_envs=5, device=dev, total_dosage=61.2, T=30, dt=1, X=200, dx=1, autoreset=True, final_obs=True, precision="fp32")
bt.reset(options={"init_condition": 0.8e5 * np.exp(-0.25 * np.linspace(0, 200, 201) ** 2)})
for _ in range(45):
    bt.step(rng.uniform(0, 0.1, 5))
for n, kwx in ((21, dict(precision="fp32")), (40, dict(precision="fp32")), (64, dict()), (64, dict(precision="fp32")), (100, dict()), (131, dict())):
    for force in ((None, "PDECG_NS_SMEM_JACOBI", "PDECG_NS_BLK_JACOBI") if n in (40, 64) else (None,)):
        if force:
            os.environ[force] = "1"
        h = 1.0 / (n - 1)
        ns = pcg.make("PDEControlGym-NavierStokes2D", num_envs=3, device=dev, boundary_condition=bc, U_ref=np.zeros((50, n, n, 2)), action_ref=np.ones(100),
                      reward_class=pcg.NSReward(), action_dim=1, maximum_pressure_iteration=23, T=0.004 * (h / 0.05) ** 2, dt=1e-3 * (h / 0.05) ** 2,
                      X=1, dx=h, Y=1, dy=h, autoreset=True, final_obs=True, **kwx)
        ns.reset(options={"init": tuple(rng.uniform(-1, 1, (3, n, n)) for _ in range(3))})
        for _ in range(6):
            ns.step(rng.uniform(2, 4, 3))
        if "precision" not in kwx:
            ns.solve_pressure(rng.uniform(-1, 1, (3, n, n)), rng.uniform(-1, 1, (3, n, n)), np.zeros((n, n)))
        if force:
            del os.environ[force]
nr = pcg.make("PDEControlGym-NeuronPDE1D", num_envs=5, device=dev, T=180, dt=5e-6, X=12e-6, dx=0.01)
nr.reset()
for _ in range(3):
    nr.step(n_steps=20)
torch.cuda.synchronize()
print("sanitize_smoke: done")

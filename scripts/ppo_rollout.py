"""BASELINE.json configs[4]: PPO-style rollout collection on TransportPDE1D + ReactionDiffusionPDE1D with the
env batch sharded over the GPUs of one box and ONE NCCL all-gather of episode statistics per rollout.

    python scripts/ppo_rollout.py [--envs 1048576] [--rollout 16] [--env rd|transport]
    torchrun --nproc-per-node 8 scripts/ppo_rollout.py ...          (global batch = --envs, sharded by rank)

The learner is out of scope: the policy is an SB3-shaped 64-64 tanh MLP in plain PyTorch that is
only evaluated (random weights); what is timed is obs -> policy -> action -> env.step for `--rollout` steps."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import pdecontrolgym_b200 as pcg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20, help="GLOBAL number of envs")
    ap.add_argument("--rollout", type=int, default=16)
    ap.add_argument("--env", default="rd", choices=["rd", "transport"])
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = pcg.shard_range(args.envs, rank, world)
    n = hi - lo
    if args.env == "rd":
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
    torch.manual_seed(0)                       # same policy replica on every rank
    policy = torch.nn.Sequential(torch.nn.Linear(npts, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                                 torch.nn.Linear(64, 1)).to(dev)
    log_std = torch.zeros(1, device=dev)
    stats = pcg.EpisodeStats(n, dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1 + rank)

    @torch.no_grad()
    def rollout():
        nonlocal obs
        for _ in range(args.rollout):
            mean = policy(obs)
            act = (mean + log_std.exp() * torch.randn(mean.shape, device=dev, generator=gen)).clamp_(-1, 1)
            obs, rew, term, trunc, _ = env.step(act)
            stats.update(rew, term, trunc)
        return stats.gather()                  # the only collective: [world, 9] doubles

    rollout()                                  # warm-up
    torch.cuda.synchronize()
    times = []
    for _ in range(args.repeats):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g = rollout()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    if rank == 0:
        ms = float(np.median(times))
        rate = args.envs * args.rollout / (ms * 1e-3)
        print(json.dumps({"config": f"PPO rollout collection, {args.env}, {args.envs} envs over {world} GPU(s), rollout {args.rollout}",
                          "n_gpus": world, "ms_per_rollout": ms, "env_steps_per_sec": rate, "grid_pt_updates_per_sec": rate * pts,
                          "stats": pcg.EpisodeStats.combine(g)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

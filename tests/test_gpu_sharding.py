"""GPU, two processes: a batch sharded over ranks by contiguous env-index ranges (`shard_range` + `env_id_offset`) reproduces
the single-batch trajectories bit for bit -- across fused auto-resets whose new initial conditions come from the on-device
Philox stream (keyed by GLOBAL env id and episode number), i.e. results do not depend on how many GPUs the batch runs on.
Rank r uses GPU r when the box has two, else both ranks share GPU 0 (the step has no collective; gloo carries the
comparison data)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _trajectory(dev, env_kind, lo, hi, steps, seed):
    import torch
    import pdecontrolgym_b200 as pcg
    if env_kind == "rd":
        eid, npts = "PDEControlGym-ReactionDiffusionPDE1D", 201
        kw = dict(T=3e-3, dt=1e-5, X=1.0, dx=5e-3, control_sample_rate=1e-3, reward_class=pcg.TunedReward1D(300, -1e3, 3e2))   # 3 steps per episode
        beta = (50 * np.cos(8 * np.arccos(np.linspace(0, 1, npts)))).astype(np.float32)
    else:
        eid, npts = "PDEControlGym-TransportPDE1D", 100
        kw = dict(T=0.04, dt=1e-4, X=1.0, dx=1e-2, control_sample_rate=0.01, reward_class=pcg.TunedReward1D(400, -1e3, 3e2))  # 4 steps per episode
        beta = (5 * np.cos(7.35 * np.arccos(np.linspace(0, 1, npts)))).astype(np.float32)
    n = hi - lo
    env = pcg.make(eid, num_envs=n, device=dev, precision="ref", normalize=True, max_control_value=20, autoreset=True,
                   env_id_offset=lo, **kw)
    obs, _ = env.reset(seed=seed, options={"init_condition": pcg.UniformScaleIC(1, 10), "beta": beta})
    rng = np.random.default_rng(99)
    acts = rng.uniform(-1, 1, (steps, 64)).astype(np.float32)          # the GLOBAL action table: env g takes column g
    out_obs, out_rew, out_done = [obs.cpu().numpy()], [], []
    for k in range(steps):
        o, r, te, tr, _ = env.step(torch.from_numpy(acts[k, lo:hi]))
        out_obs.append(o.cpu().numpy())
        out_rew.append(r.cpu().numpy())
        out_done.append((te | tr).cpu().numpy())
    return np.stack(out_obs), np.stack(out_rew), np.stack(out_done)


def _worker(rank, world, port, env_kind, n_global, steps, seed, ret):
    import torch
    import torch.distributed as dist
    import pdecontrolgym_b200 as pcg
    os.environ["PDECG_BC1D_HALFWARP_MIN"] = "0"
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    lo, hi = pcg.shard_range(n_global, rank, world)
    obs, rew, done = _trajectory(dev, env_kind, lo, hi, steps, seed)
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, obs, rew, done))
    if rank == 0:
        full_obs, full_rew, full_done = _trajectory(dev, env_kind, 0, n_global, steps, seed)      # the un-sharded batch
        ok = True
        for lo_r, hi_r, o, r, d in parts:
            ok &= np.array_equal(o, full_obs[:, lo_r:hi_r]) and np.array_equal(r, full_rew[:, lo_r:hi_r]) and np.array_equal(d, full_done[:, lo_r:hi_r])
        ret["ok"] = bool(ok)
        ret["resets"] = int(full_done.sum())
        ret["distinct_ics"] = int(len(np.unique(full_obs[:, :, 0])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("env_kind", ["rd", "transport"])
def test_two_rank_shards_reproduce_the_single_batch(cuda_device, env_kind):
    import torch.multiprocessing as mp
    n_global, steps = 37, 9                      # uneven shards (19 + 18); every env restarts at least twice
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, env_kind, n_global, steps, 7, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=300)
            assert p.exitcode == 0
        assert ret["ok"], "sharded trajectories differ from the single-batch run"
        assert ret["resets"] >= 2 * n_global and ret["distinct_ics"] > n_global     # fresh Philox draws per episode were compared

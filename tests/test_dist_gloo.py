"""CPU, world_size 2 over gloo: env-index sharding and the per-rollout statistics all-gather (the only
collective of the N>1 path).  Sharding invariance of the CUDA kernels themselves is in the GPU tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pdecontrolgym_b200.dist import STAT_FIELDS, EpisodeStats, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 65536, 2 ** 20 + 3):
        for w in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _episode_stream(n_envs, steps, seed):
    rng = np.random.default_rng(seed)
    rew = rng.standard_normal((steps, n_envs))
    term = rng.uniform(size=(steps, n_envs)) < 0.05
    trunc = rng.uniform(size=(steps, n_envs)) < 0.02
    return rew, term, trunc


def _run_stats(rew, term, trunc, lo, hi):
    st = EpisodeStats(hi - lo, "cpu")
    for k in range(rew.shape[0]):
        st.update(torch.from_numpy(rew[k, lo:hi]), torch.from_numpy(term[k, lo:hi]), torch.from_numpy(trunc[k, lo:hi]))
    return st


def _worker(rank, world, port, n_envs, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rew, term, trunc = _episode_stream(n_envs, steps, 42)
    lo, hi = shard_range(n_envs, rank, world)
    st = _run_stats(rew, term, trunc, lo, hi)
    g = st.gather()
    assert g.shape == (world, len(STAT_FIELDS))
    if rank == 0:
        q.put(EpisodeStats.combine(g))
    dist.barrier()
    dist.destroy_process_group()


def test_stats_allgather_world2_matches_single_process():
    n_envs, steps = 37, 200
    rew, term, trunc = _episode_stream(n_envs, steps, 42)
    single = EpisodeStats.combine(_run_stats(rew, term, trunc, 0, n_envs).gather())
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_envs, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    multi = q.get()
    assert multi["episodes"] == single["episodes"] > 0
    for k in ("truncations", "terminations", "env_steps"):
        assert multi[k] == single[k]
    for k in ("return_mean", "return_std", "length_mean", "return_min", "return_max"):
        assert abs(multi[k] - single[k]) <= 1e-12 * max(1.0, abs(single[k])), k

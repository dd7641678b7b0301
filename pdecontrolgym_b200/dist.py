"""Multi-GPU plumbing: one process per GPU (torchrun), the env batch sharded by contiguous index ranges,
NO collective inside `step`/`reset`.  The only exchange is a small per-rollout statistics vector that
every rank all-gathers for a centralised learner / logger (NCCL over NVLink on GPUs; gloo in the CPU
tests).  The reference has no counterpart -- it never builds a vector env."""
from __future__ import annotations

import torch
import torch.distributed as dist

STAT_FIELDS = ("episodes", "return_sum", "return_sq_sum", "length_sum", "truncations", "terminations",
               "return_min", "return_max", "env_steps")


def shard_range(global_n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of the global env index owned by `rank` (sizes differ by at most 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(global_n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class EpisodeStats:
    """Per-env running return / length and per-rank episode statistics, kept on the env's device.
    `update` is a handful of elementwise torch ops on [num_envs] tensors -- it never synchronises."""

    def __init__(self, num_envs: int, device):
        self.device = torch.device(device)
        self.ret = torch.zeros(num_envs, dtype=torch.float64, device=self.device)
        self.length = torch.zeros(num_envs, dtype=torch.int64, device=self.device)
        self.reset_totals()

    def reset_totals(self):
        d = self.device
        self.tot = torch.zeros(len(STAT_FIELDS), dtype=torch.float64, device=d)
        self.tot[STAT_FIELDS.index("return_min")] = float("inf")
        self.tot[STAT_FIELDS.index("return_max")] = float("-inf")

    def update(self, reward, terminated, truncated):
        """Mask arithmetic only (no boolean indexing, no data-dependent shapes): nothing here synchronises with the
        host, and the whole update can be captured in a CUDA graph."""
        if reward.numel() == 0:
            return
        done = terminated | truncated
        self.ret += reward.to(torch.float64)
        self.length += 1
        zero = torch.zeros((), dtype=torch.float64, device=self.device)
        fin = torch.where(done, self.ret, zero)
        t = self.tot
        t[0] += done.sum()
        t[1] += fin.sum()
        t[2] += (fin * fin).sum()
        t[3] += torch.where(done, self.length, torch.zeros_like(self.length)).sum()
        t[4] += truncated.sum()
        t[5] += (terminated & ~truncated).sum()
        inf = torch.full((), float("inf"), dtype=torch.float64, device=self.device)
        t[6] = torch.minimum(t[6], torch.where(done, self.ret, inf).min())
        t[7] = torch.maximum(t[7], torch.where(done, self.ret, -inf).max())
        t[8] += reward.numel()
        self.ret = torch.where(done, torch.zeros_like(self.ret), self.ret)
        self.length = torch.where(done, torch.zeros_like(self.length), self.length)

    def gather(self):
        """All ranks' totals, shape [world, len(STAT_FIELDS)] (one all_gather; [1, F] without a process group)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            out = torch.empty(dist.get_world_size() * self.tot.numel(), dtype=self.tot.dtype, device=self.device)
            dist.all_gather_into_tensor(out, self.tot.contiguous())
            return out.view(dist.get_world_size(), -1)
        return self.tot.view(1, -1).clone()

    @staticmethod
    def combine(gathered):
        """Global statistics dict from `gather()`'s output."""
        g = gathered.to(torch.float64)
        i = STAT_FIELDS.index
        n = g[:, i("episodes")].sum()
        mean = g[:, i("return_sum")].sum() / n.clamp(min=1)
        var = g[:, i("return_sq_sum")].sum() / n.clamp(min=1) - mean * mean
        return {
            "episodes": int(n.item()), "return_mean": float(mean.item()), "return_std": float(var.clamp(min=0).sqrt().item()),
            "length_mean": float((g[:, i("length_sum")].sum() / n.clamp(min=1)).item()),
            "truncations": int(g[:, i("truncations")].sum().item()), "terminations": int(g[:, i("terminations")].sum().item()),
            "return_min": float(g[:, i("return_min")].min().item()), "return_max": float(g[:, i("return_max")].max().item()),
            "env_steps": int(g[:, i("env_steps")].sum().item()),
        }

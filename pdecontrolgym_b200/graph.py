"""CUDA-graph replay of a whole env step.

A batched `env.step()` is one integrator kernel plus a handful of small launches (same-step auto-reset, flag /
reward / observation copies).  For small batches (e.g. the 16,384-env Traffic-ARZ / BrainTumor configs) the
per-launch host cost dominates the ~20 us of GPU work, so the step is captured once into a CUDA graph and replayed:
one `cudaGraphLaunch` per step, no Python between the kernels.

    env.reset(...)
    g = GraphedStep(env, torch.zeros(env.num_envs, device=env.device))   # 3 warm-up steps + capture
    env.reset(...)                                                      # capture advanced the state: reset again
    obs, reward, terminated, truncated, info = g.step(actions)          # outputs are STATIC tensors, overwritten by
                                                                        # the next g.step(); clone what you keep
"""
from __future__ import annotations

import torch


class GraphedStep:
    def __init__(self, env, example_action, warmup: int = 3):
        self.env = env
        dev = env.device if hasattr(env, "device") else env.env.device
        self.static_action = torch.as_tensor(example_action).to(dev).clone().contiguous()
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):            # lazy module loads, allocator warm-up, cudaFuncSetAttribute calls
                env.step(self.static_action)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = env.step(self.static_action)

    def step(self, action):
        self.static_action.copy_(torch.as_tensor(action).reshape(self.static_action.shape), non_blocking=True)
        self.graph.replay()
        return self.out

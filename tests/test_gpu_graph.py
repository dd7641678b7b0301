"""GPU: CUDA-graph replay of env.step() reproduces eager stepping bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(make, reset, N, act_fn, steps, device):
    import torch
    import pdecontrolgym_b200 as pcg
    eager, graphed = make(), make()
    reset(eager)
    reset(graphed)
    g = pcg.GraphedStep(graphed, act_fn(0))
    reset(graphed)
    for k in range(steps):
        a = act_fn(k + 1)
        o1, r1, te1, tr1, _ = eager.step(a)
        o2, r2, te2, tr2, _ = g.step(a)
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(te1, te2) and torch.equal(tr1, tr2), f"step {k}"


def test_graph_rd_with_autoreset(cuda_device):
    import torch
    import pdecontrolgym_b200 as pcg
    N = 50
    kw = dict(T=0.004, dt=1e-5, X=1, dx=5e-3, control_sample_rate=1e-3, normalize=True, reward_class=pcg.TunedReward1D(400), autoreset=True)
    rng = np.random.default_rng(0)
    acts = torch.from_numpy(rng.uniform(-1, 1, (20, N)).astype(np.float32)).to(cuda_device)
    ic = rng.uniform(1, 3, (N, 201))
    _pair(lambda: pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=N, device=cuda_device, **kw),
          lambda e: e.reset(options={"init_condition": ic, "beta": np.linspace(0, 5, 201)}), N, lambda k: acts[k], 12, cuda_device)


def test_graph_arz_and_brain_tumor(cuda_device):
    import torch
    import pdecontrolgym_b200 as pcg
    N = 33
    rng = np.random.default_rng(1)
    acts = torch.from_numpy(rng.uniform(1.0, 1.4, (40, N, 1))).to(cuda_device)
    _pair(lambda: pcg.make("PDEControlGym-TrafficPDE1D", num_envs=N, device=cuda_device, simulation_type="outlet", T=2, dt=0.25, X=500, dx=10, autoreset=True),
          lambda e: (e.reset(), e.time_index.zero_()), N, lambda k: acts[k], 30, cuda_device)   # the reference's reset() keeps the clock
    xs = np.linspace(0, 200, 201)
    ic = 0.8e5 * np.exp(-0.25 * xs ** 2)
    doses = torch.from_numpy(rng.uniform(0, 0.1, (40, N))).to(cuda_device)
    _pair(lambda: pcg.make("PDEControlGym-BrainTumor1D", num_envs=N, device=cuda_device, total_dosage=61.2, T=600, dt=1, X=200, dx=1),
          lambda e: e.reset(options={"init_condition": ic}), N, lambda k: doses[k], 30, cuda_device)

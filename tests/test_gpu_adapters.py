"""GPU: vector-env protocol adapters and the ARZ backstepping driver."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_adapters_shapes_and_semantics(cuda_device):
    import pdecontrolgym_b200 as pcg
    from pdecontrolgym_b200.adapters import GymnasiumVectorAdapter, SB3VecEnvAdapter
    kw = dict(T=0.002, dt=1e-5, X=1, dx=5e-3, control_sample_rate=1e-3, normalize=True, reward_class=pcg.TunedReward1D(200))
    env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=6, device=cuda_device, autoreset=True, **kw)
    opts = {"init_condition": pcg.UniformScaleIC(1, 10), "beta": np.zeros(201)}
    g = GymnasiumVectorAdapter(env)
    obs, info = g.reset(seed=3, options=opts)
    assert isinstance(obs, np.ndarray) and obs.shape == (6, 201) and obs.dtype == np.float32
    obs, r, te, tr, info = g.step(np.zeros((6, 1), dtype=np.float32))
    assert r.shape == (6,) and te.dtype == bool and not te.any()
    obs, r, te, tr, info = g.step(np.zeros((6, 1), dtype=np.float32))
    assert te.all()
    s = SB3VecEnvAdapter(env)
    obs = s.reset()
    s.step_async(np.zeros((6, 1), dtype=np.float32))
    obs, r, dones, infos = s.step_wait()
    assert obs.shape == (6, 201) and r.dtype == np.float32 and len(infos) == 6 and not dones.any()
    obs, r, dones, infos = s.step(np.zeros((6, 1), dtype=np.float32))
    assert dones.all() and not infos[0]["TimeLimit.truncated"]
    with pytest.raises(ValueError):
        SB3VecEnvAdapter(pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=2, device=cuda_device, **kw))


def test_arz_backstepping_drives_to_steady_state(cuda_device):
    """Closed loop with the tutorial's outlet law: the reward (minus distance to the steady state) improves
    markedly over the open-loop run, for every env of the batch (Trafficarz1d_tutorial.rst:153-184)."""
    import torch
    import pdecontrolgym_b200 as pcg
    from pdecontrolgym_b200.controllers import TrafficARZBackstepping
    kw = dict(T=240, dt=0.25, X=500, dx=10, simulation_type="outlet", num_envs=4, device=cuda_device)
    closed, opened = pcg.make("PDEControlGym-TrafficPDE1D", **kw), pcg.make("PDEControlGym-TrafficPDE1D", **kw)
    ctl = TrafficARZBackstepping(closed)
    closed.reset()
    opened.reset()
    for k in range(959):
        oc, rc, *_ = closed.step(ctl())
        oo, ro, *_ = opened.step(torch.full((4, 1), 1.2, dtype=torch.float64))
    assert torch.isfinite(rc).all() and (rc > ro).all() and (rc > -0.2).all(), (rc, ro)

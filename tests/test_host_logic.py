"""CPU: host-side logic that needs no GPU -- registry ids, reward descriptors, constructor validation
messages (same text as the reference), kernel-variant selection rules."""
import math

import numpy as np
import pytest

import pdecontrolgym_b200 as pcg


def test_registry_has_the_five_reference_ids():
    # pde_control_gym/__init__.py:4,8,12,13,17
    assert sorted(pcg.REGISTRY) == sorted([
        "PDEControlGym-TransportPDE1D", "PDEControlGym-ReactionDiffusionPDE1D", "PDEControlGym-BrainTumor1D",
        "PDEControlGym-TrafficPDE1D", "PDEControlGym-NavierStokes2D"])
    with pytest.raises(KeyError):
        pcg.make("PDEControlGym-NeuronPDE1D")


def test_reward_descriptors_mirror_reference_signatures():
    r = pcg.TunedReward1D(50000, -1e3, 3e2)
    assert (r.nt, r.truncate_penalty, r.terminate_reward) == (50000, -1e3, 3e2)
    assert pcg.TunedReward1D(10).truncate_penalty == -1e-4 and pcg.TunedReward1D(10).terminate_reward == 1e2
    with pytest.raises(Exception, match="Number of simulation steps"):
        pcg.TunedReward1D(None)
    with pytest.raises(Exception, match="Number of simulation steps"):
        pcg.NormReward()
    n = pcg.NormReward(100, "inf", "t-horizon", t_horizon_length=3)
    assert n._ord == 3 and n._horizon == 2 and n.norm_coeff == 1.0
    assert pcg.NormReward(100, norm=math.inf)._ord == 3 and pcg.NormReward(100, norm=1)._ord == 1
    with pytest.raises(ValueError):
        pcg.NormReward(100, norm="3")
    with pytest.raises(ValueError):
        pcg.NormReward(100, horizon="spatial")
    assert pcg.NSReward().gamma == 0.1
    assert issubclass(pcg.NormReward, pcg.BaseReward) and issubclass(pcg.BrainTumorReward, pcg.BaseReward)
    with pytest.raises(NotImplementedError):
        pcg.TunedReward1D(10).reward(None, 0, False, False, 0.0)


def test_constructor_validation_precedes_device_use():
    kw = dict(T=1, dt=1e-4, X=1, dx=1e-2, num_envs=1)
    with pytest.raises(ValueError, match="precision"):
        pcg.make("PDEControlGym-TransportPDE1D", precision="fp16", reward_class=pcg.TunedReward1D(10000), **kw)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        pcg.make("PDEControlGym-TransportPDE1D", device="cpu", reward_class=pcg.TunedReward1D(10000), **kw)
    with pytest.raises(NotImplementedError, match="inlet-train"):
        pcg.make("PDEControlGym-TrafficPDE1D", simulation_type="inlet-train", T=240, dt=0.25, X=500, dx=10)
    with pytest.raises(ValueError, match="Invalid simulation type"):
        pcg.make("PDEControlGym-TrafficPDE1D", simulation_type="x", T=240, dt=0.25, X=500, dx=10)
    with pytest.raises(TypeError):
        pcg.make("PDEControlGym-NavierStokes2D", T=0.2, dt=1e-3, X=1, dx=0.05, Y=1, dy=0.05, reward_class=None)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        pcg.make("PDEControlGym-BrainTumor1D", T=600, dt=1, X=200, dx=1, device="cpu", total_dosage=61.2)


def test_traffic_static_helpers_match_reference_formulas():
    T = pcg.TrafficPDE1D
    assert T.Veq(40, 0.16, 0.12) == 10.000000000000002 or T.Veq(40, 0.16, 0.12) == 40 * (1 - 0.12 / 0.16)
    rho, y = np.array([0.1, 0.12]), np.array([0.3, -0.2])
    assert np.array_equal(T.F_r(40, 0.16, rho, y), y + rho * (40 * (1 - rho / 0.16)))
    assert np.array_equal(T.F_y(40, 0.16, rho, y), y * (y / rho + 40 * (1 - rho / 0.16)))


def test_host_buffer_argument_validation():
    """`step_host` arguments are checked on the host before any CUDA call (pdecontrolgym_b200._lib.host_ptr)."""
    import torch
    from pdecontrolgym_b200 import _lib
    assert _lib.host_ptr(None, torch.float64, 4, "x") is None
    t = torch.zeros(4, dtype=torch.float64)
    assert _lib.host_ptr(t, torch.float64, 4, "x") == t.data_ptr()
    for bad in (torch.zeros(4, dtype=torch.float32), torch.zeros(5, dtype=torch.float64), torch.zeros(2, 4, dtype=torch.float64)[:, ::2]):
        with pytest.raises(ValueError, match="contiguous CPU"):
            _lib.host_ptr(bad, torch.float64, 4, "x")
    assert issubclass(pcg.PdecgError, RuntimeError) and pcg.PdecgError(-2, "msg").code == -2


def test_adjoint_wall_zeroing_matches_the_example():
    """controllers.NavierStokesAdjoint._zero_walls == apply_boundary of examples/NavierStokes/NS2Doptimization.py:55-60."""
    import torch
    from pdecontrolgym_b200.controllers import NavierStokesAdjoint
    rng = np.random.default_rng(0)
    a1, a2 = rng.uniform(-1, 1, (3, 7, 7)), rng.uniform(-1, 1, (3, 7, 7))
    t1, t2 = NavierStokesAdjoint._zero_walls(torch.from_numpy(a1.copy()), torch.from_numpy(a2.copy()))
    for a in (a1, a2):
        a[:, :, [-1, 0]] = 0.
        a[:, [-1, 0], :] = 0.
    assert np.array_equal(t1.numpy(), a1) and np.array_equal(t2.numpy(), a2)


def test_batched_helpers_match_the_reference_formulas():
    """Batched central_difference / laplace (navier_stokes2D.py:9-22) on CPU tensors vs the slice expressions."""
    import torch
    f = np.random.default_rng(1).uniform(-1, 1, (2, 9, 9))
    t = torch.from_numpy(f)
    want = np.zeros_like(f)
    want[:, 1:-1, 1:-1] = (f[:, 1:-1, 2:] - f[:, 1:-1, 0:-2]) / (2 * 0.05)
    np.testing.assert_allclose(pcg.central_difference(t, "x", 0.05).numpy(), want, rtol=1e-15, atol=0)
    want = np.zeros_like(f)
    want[:, 1:-1, 1:-1] = (f[:, 2:, 1:-1] - f[:, 0:-2, 1:-1]) / (2 * 0.05)
    np.testing.assert_allclose(pcg.central_difference(t, "y", 0.05).numpy(), want, rtol=1e-15, atol=0)
    want = np.zeros_like(f)
    want[:, 1:-1, 1:-1] = (f[:, 1:-1, 0:-2] + f[:, 0:-2, 1:-1] - 4 * f[:, 1:-1, 1:-1] + f[:, 1:-1, 2:] + f[:, 2:, 1:-1]) / (0.05 * 0.05)
    np.testing.assert_allclose(pcg.laplace(t, 0.05, 0.05).numpy(), want, rtol=1e-14, atol=0)


class _FakeVecEnv:
    """Minimal stand-in with the vector-env surface of this package (CPU tensors) for the adapter logic tests."""

    def __init__(self, n=3, autoreset=True):
        import torch
        self.num_envs, self.autoreset, self.t = n, autoreset, 0
        self.single_observation_space = self.single_action_space = self.observation_space = self.action_space = None
        self._torch = torch

    def reset(self, seed=None, options=None):
        self.t = 0
        return self._torch.zeros(self.num_envs, 4), {}

    def step(self, actions):
        self.t += 1
        torch = self._torch
        obs = torch.full((self.num_envs, 4), float(self.t))
        term = torch.tensor([self.t % 3 == 0, False, False][:self.num_envs])
        trunc = torch.tensor([False, self.t % 2 == 0, False][:self.num_envs])
        return obs, torch.arange(self.num_envs, dtype=torch.float64), term, trunc, {"valid": torch.ones(self.num_envs, dtype=torch.bool)}

    def close(self):
        self.closed = True


def test_protocol_adapters_on_a_fake_env():
    from pdecontrolgym_b200.adapters import GymnasiumVectorAdapter, SB3VecEnvAdapter, TrajectoryRecorder
    g = GymnasiumVectorAdapter(_FakeVecEnv())
    obs, info = g.reset(seed=1)
    assert isinstance(obs, np.ndarray) and obs.shape == (3, 4) and info == {}
    obs, rew, term, trunc, info = g.step(np.zeros(3))
    assert all(isinstance(x, np.ndarray) for x in (obs, rew, term, trunc, info["valid"])) and term.dtype == bool
    with pytest.raises(ValueError, match="autoreset=True"):
        SB3VecEnvAdapter(_FakeVecEnv(autoreset=False))
    s = SB3VecEnvAdapter(_FakeVecEnv())
    assert s.reset().shape == (3, 4)
    s.step_async(np.zeros(3))
    s.step_wait()
    obs, rew, dones, infos = s.step(np.zeros(3))                  # t == 2: env 1 truncated
    assert rew.dtype == np.float32 and dones.tolist() == [False, True, False]
    assert [i["TimeLimit.truncated"] for i in infos] == [False, True, False]
    obs, rew, dones, infos = s.step(np.zeros(3))                  # t == 3: env 0 terminated (not a time-limit truncation)
    assert dones.tolist() == [True, False, False] and not infos[0]["TimeLimit.truncated"]
    assert s.env_is_wrapped(object) == [False] * 3 and s.seed(7) == [7, 7, 7]
    rec = TrajectoryRecorder(_FakeVecEnv(), max_steps=2)
    rec.reset()
    for _ in range(4):                                            # only the first max_steps steps are kept
        rec.step(None)
    h = rec.history()
    assert h.shape == (3, 3, 4) and h[:, 0, 0].tolist() == [0.0, 1.0, 2.0]

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
        "PDEControlGym-TrafficPDE1D", "PDEControlGym-NavierStokes2D",
        "PDEControlGym-NeuronPDE1D"])          # not registered upstream (neuron_env.py is not exported); offered here
    with pytest.raises(KeyError):
        pcg.make("PDEControlGym-NoSuchEnv")


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
    """Minimal stand-in with the vector-env surface of this package (CPU tensors) for the adapter logic tests: env i counts
    its own steps, observation = (episode step, env index), env 0 terminates every 3rd step, env 1 truncates every 2nd."""

    def __init__(self, n=3, autoreset=True, final_obs=True):
        import torch
        self.num_envs, self.autoreset = n, autoreset
        self.single_observation_space = self.single_action_space = self.observation_space = self.action_space = None
        self._torch = torch
        self.t = torch.zeros(n, dtype=torch.int64)
        self._final = torch.zeros(n, 4) if (autoreset and final_obs) else None

    def _obs(self):
        return self._torch.stack([self.t.double(), self._torch.arange(self.num_envs).double(),
                                  self._torch.zeros(self.num_envs).double(), self._torch.ones(self.num_envs).double()], dim=1).float()

    def reset(self, seed=None, options=None):
        mask = (options or {}).get("mask")
        if mask is None:
            self.t.zero_()
        else:
            self.t[self._torch.as_tensor(mask).bool()] = 0
        return self._obs(), {}

    def step(self, actions):
        torch = self._torch
        period = torch.tensor([3, 2, 10 ** 9][:self.num_envs])
        finished = (self.t > 0) & (self.t % period == 0)          # stepping a finished env leaves it alone (like the CUDA envs)
        self.t = torch.where(finished, self.t, self.t + 1)
        hit = self.t % period == 0
        term = hit & (torch.arange(self.num_envs) == 0)
        trunc = hit & (torch.arange(self.num_envs) == 1)
        info = {"valid": torch.ones(self.num_envs, dtype=torch.bool)}
        obs = self._obs()
        if self.autoreset:
            done = term | trunc
            if self._final is not None:
                self._final = torch.where(done[:, None], obs, self._final)
                info["final_obs"] = self._final
            self.t = torch.where(done, torch.zeros_like(self.t), self.t)
            obs = self._obs()
        return obs, torch.arange(self.num_envs, dtype=torch.float64) + 1, term, trunc, info

    def close(self):
        self.closed = True


def test_protocol_adapters_on_a_fake_env():
    from pdecontrolgym_b200.adapters import GymnasiumVectorAdapter, SB3VecEnvAdapter, TrajectoryRecorder
    from tests.vecenv_protocol import drive_gymnasium, drive_sb3
    act = lambda k: np.zeros(3)  # noqa: E731
    # gymnasium, same-step auto-reset: terminal observations in infos["final_obs"], new episode's first row in obs
    log = drive_gymnasium(GymnasiumVectorAdapter(_FakeVecEnv()), act, 7, (4,))
    obs, rew, term, trunc, infos = log[1]                          # step 2: env 1 truncated
    assert trunc.tolist() == [False, True, False] and infos["final_obs"][1, 0] == 2 and obs[1, 0] == 0 and obs[0, 0] == 2
    obs, rew, term, trunc, infos = log[2]                          # step 3: env 0 terminated
    assert term.tolist() == [True, False, False] and infos["final_obs"][0, 0] == 3 and obs[0, 0] == 0
    assert isinstance(infos["valid"], np.ndarray)
    # gymnasium, next-step auto-reset (driven by the adapter on an env without auto-reset)
    with pytest.raises(ValueError, match="autoreset=False"):
        GymnasiumVectorAdapter(_FakeVecEnv(), autoreset_mode="next_step")
    log = drive_gymnasium(GymnasiumVectorAdapter(_FakeVecEnv(autoreset=False), autoreset_mode="next_step"), act, 9, (4,))
    assert log[1][3].tolist() == [False, True, False] and log[1][0][1, 0] == 2        # terminal row returned by the step itself
    assert log[2][0][1, 0] == 0 and log[2][1][1] == 0 and log[2][2].tolist() == [True, False, False]
    assert log[3][0][0, 0] == 0 and log[3][0][1, 0] == 1                               # env 0 reset, env 1 stepping again
    # stable-baselines3
    with pytest.raises(ValueError, match="autoreset=True"):
        SB3VecEnvAdapter(_FakeVecEnv(autoreset=False))
    log = drive_sb3(SB3VecEnvAdapter(_FakeVecEnv()), act, 7, (4,))
    obs, rew, dones, infos = log[1]
    assert dones.tolist() == [False, True, False] and infos[1]["TimeLimit.truncated"] and infos[1]["terminal_observation"][0] == 2
    obs, rew, dones, infos = log[2]
    assert dones.tolist() == [True, False, False] and not infos[0]["TimeLimit.truncated"] and infos[0]["terminal_observation"][0] == 3
    s = SB3VecEnvAdapter(_FakeVecEnv())
    assert s.seed(7) == [7, 7, 7]
    rec = TrajectoryRecorder(_FakeVecEnv(autoreset=False), max_steps=2, extras=("t",))
    rec.reset()
    for _ in range(4):                                            # only the first max_steps steps are kept
        rec.step(None)
    h = rec.history()
    assert h.shape == (3, 3, 4) and h[:, 2, 0].tolist() == [0.0, 1.0, 2.0]
    assert rec.extra_history("t")[:, 2].tolist() == [0, 1, 2]

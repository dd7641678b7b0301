"""Reward classes with the reference's names and constructor signatures.

In the reference these objects *compute* the reward from the trajectory buffer on the host
(`reward(uVec, time_index, terminate, truncate, action)`, base_reward.py:14).  Here they are
descriptors: the env's CUDA kernel evaluates the reward in the same launch as the PDE update, from
O(1) carried scalars instead of the `u[nt, nx]` history.  Arbitrary Python `BaseReward` subclasses
cannot run inside a kernel and are rejected with a clear error.
"""
import math


class BaseReward:
    """Mirror of rewards/base_reward.py:5-32."""

    def reward(self, uVec=None, time_index=None, terminate=None, truncate=None, action=None):
        raise NotImplementedError(
            "pdecontrolgym_b200 evaluates rewards inside the env kernel; custom Python rewards are not "
            "supported on the batched path (compute them from the returned observations instead).")

    def reset(self):
        pass


class TunedReward1D(BaseReward):
    """rewards/tuned_reward_1d.py:17-40."""

    def __init__(self, nt: int, truncate_penalty: float = -1e-4, terminate_reward: float = 1e2):
        if nt is None:
            raise Exception("Number of simulation steps must be specified in the NormReward class.")
        self.nt = nt
        self.truncate_penalty = truncate_penalty
        self.terminate_reward = terminate_reward


class NormReward(BaseReward):
    """rewards/norm_reward.py:19-73 with the repaired semantics documented in DESIGN.md (the
    reference class cannot run as shipped).  `norm` accepts "1" | "2" | "inf" (or 1, 2, math.inf);
    `norm_coeff` (an undefined global in the reference) is an explicit argument, default 1.0."""

    def __init__(self, nt: int = None, norm: str = "2", horizon: str = "temporal",
                 truncate_penalty: float = -1e-4, terminate_reward: float = 1e2,
                 t_horizon_length: int = 5, *extras, norm_coeff: float = 1.0):
        if nt is None:
            raise Exception("Number of simulation steps must be specified in the NormReward class.")
        ords = {"1": 1, "2": 2, "inf": 3, 1: 1, 2: 2, math.inf: 3}
        if norm not in ords:
            raise ValueError(f"NormReward: norm must be '1', '2' or 'inf' (got {norm!r})")
        horizons = {"temporal": 0, "differential": 1, "t-horizon": 2}
        if horizon not in horizons:
            raise ValueError(f"NormReward: horizon must be one of {sorted(horizons)} (got {horizon!r})")
        self.nt = nt
        self.norm = norm
        self.horizon = horizon
        self.truncate_penalty = truncate_penalty
        self.terminate_reward = terminate_reward
        self.t_horizon_length = t_horizon_length
        self.norm_coeff = norm_coeff
        self._ord = ords[norm]
        self._horizon = horizons[horizon]


class NSReward(BaseReward):
    """rewards/ns_reward.py:15-28."""

    def __init__(self, gamma: float = 0.1):
        self.gamma = gamma


class TrafficARZReward(BaseReward):
    """rewards/traffic_arz_reward.py:5-22."""


class BrainTumorReward(BaseReward):
    """rewards/brain_tumor_reward.py:5-73."""

"""pdecontrolgym_b200 -- B200-native batched PDE-control environments.

Drop-in, *vector* counterpart of lukebhan/PDEControlGym's environment dynamics: same gymnasium ids,
constructor keywords and reward-class names; `reset` / `step` act on `num_envs` lock-stepped
instances whose state lives on the GPU and is advanced by hand-written sm_100a kernels
(`csrc/*.cu`) through the C ABI in `include/pdecg.h`.  There is no CPU path.

    import pdecontrolgym_b200 as pcg
    env = pcg.make("PDEControlGym-ReactionDiffusionPDE1D", num_envs=65536, device="cuda:0",
                   precision="fp64", **params)        # params: the reference's kwargs dict
    obs, info = env.reset()
    obs, reward, terminated, truncated, info = env.step(actions)
"""
from .rewards import (BaseReward, BrainTumorReward, NormReward, NSReward, TrafficARZReward,
                      TunedReward1D)
from .envs_bc1d import ReactionDiffusionPDE1D, TransportPDE1D, UniformScaleIC
from .envs_brain_tumor import BrainTumor1D, TherapyWrapper
from .envs_traffic_arz import TrafficPDE1D
from .envs_navier_stokes import NavierStokes2D, central_difference, laplace
from .envs_neuron import NeuronPDE1D
from .dist import EpisodeStats, shard_range
from . import adapters, controllers
from .graph import GraphedStep
from ._lib import PdecgError

__version__ = "0.1.0"

# id -> class: the five ids of pde_control_gym/__init__.py:3-18 (the reference's own registration
# file has a syntax error; all five are registered here)
REGISTRY = {
    "PDEControlGym-TransportPDE1D": TransportPDE1D,
    "PDEControlGym-ReactionDiffusionPDE1D": ReactionDiffusionPDE1D,
    "PDEControlGym-BrainTumor1D": BrainTumor1D,
    "PDEControlGym-TrafficPDE1D": TrafficPDE1D,
    "PDEControlGym-NavierStokes2D": NavierStokes2D,
    # not registered by the reference (neuron_env.py is not even exported there); offered under the id its docs imply
    "PDEControlGym-NeuronPDE1D": NeuronPDE1D,
}


def make(id: str, num_envs: int = 1, device="cuda", precision: str = "ref", **kwargs):
    """`gym.make(id, **kwargs)` for a batch: returns a vector env of `num_envs` instances."""
    if id not in REGISTRY:
        raise KeyError(f"unknown environment id {id!r}; known: {sorted(REGISTRY)}")
    return REGISTRY[id](num_envs=num_envs, device=device, precision=precision, **kwargs)


make_vec = make

__all__ = ["make", "make_vec", "REGISTRY", "TransportPDE1D", "ReactionDiffusionPDE1D", "BrainTumor1D", "TherapyWrapper", "TrafficPDE1D", "NavierStokes2D", "NeuronPDE1D", "central_difference", "laplace", "EpisodeStats", "shard_range", "UniformScaleIC", "GraphedStep", "PdecgError", "BaseReward",
           "NormReward", "TunedReward1D", "NSReward", "TrafficARZReward", "BrainTumorReward"]

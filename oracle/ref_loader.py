"""Loader for the UNMODIFIED reference (lukebhan/PDEControlGym) -- TEST INFRASTRUCTURE ONLY.

The reference is pure Python + numpy but cannot be imported as shipped:
  * `pde_control_gym/__init__.py:11-14` is a botched merge (SyntaxError);
  * `gymnasium` / `matplotlib` are not installed in this image.
This module injects the stubs in `oracle/_stubs` and pre-registers empty package objects so
the env / reward *sub-modules* are imported straight from `/root/reference` without touching
the broken `__init__` files.  No reference source is copied into this repository.

Where the reference comes from (first hit wins): `$PDECG_REFERENCE_ROOT`, `/root/reference` (build container
only), `oracle/_ref/` -- the verbatim, git-ignored copy staged by `oracle/make_ref.py` at build() time, which
travels to the GPU box with the snapshot.  It is used
  (1) by `tests/golden/gen_golden.py` to produce the committed golden vectors,
  (2) by `tests/test_oracle_vs_reference.py`, which is skipped when no tree is found, and
  (3) by `bench.py` as the CPU arm (`--impl reference`, `cpu_baseline.kind == "reference"`).
Nothing under `pdecontrolgym_b200/` may import this file.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "_stubs")


def _find_root():
    cands = [os.environ.get("PDECG_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "pde_control_gym", "src")):
            return c
    return cands[0] or "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_kind() -> str:
    """'live' (the read-only checkout in the build container) or 'staged' (oracle/_ref copy)."""
    return "staged" if os.path.abspath(REFERENCE_ROOT) == os.path.join(_HERE, "_ref") else "live"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pde_control_gym", "src"))


_loaded = None


def load_reference():
    """Return a namespace with the reference classes (imports them on first use)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    try:
        import gymnasium  # noqa: F401  (real one wins if it is ever installed)
    except ImportError:
        sys.path.insert(0, _STUBS)
    try:
        import matplotlib.pyplot  # noqa: F401
    except ImportError:
        if _STUBS not in sys.path:
            sys.path.insert(0, _STUBS)
    root = os.path.join(REFERENCE_ROOT, "pde_control_gym")
    for name, path in (
        ("pde_control_gym", root),
        ("pde_control_gym.src", os.path.join(root, "src")),
    ):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = [path]
            sys.modules[name] = mod
    ns = types.SimpleNamespace()
    imp = importlib.import_module
    ns.TransportPDE1D = imp("pde_control_gym.src.environments1d.hyperbolic").TransportPDE1D
    ns.ReactionDiffusionPDE1D = imp("pde_control_gym.src.environments1d.parabolic").ReactionDiffusionPDE1D
    ns.TrafficPDE1D = imp("pde_control_gym.src.environments1d.traffic_arz_env").TrafficPDE1D
    bt = imp("pde_control_gym.src.environments1d.brain_tumor_env")
    ns.BrainTumor1D, ns.TherapyWrapper = bt.BrainTumor1D, bt.TherapyWrapper
    ns2d = imp("pde_control_gym.src.environments2d.navier_stokes2D")
    ns.NavierStokes2D = ns2d.NavierStokes2D
    ns.central_difference, ns.laplace = ns2d.central_difference, ns2d.laplace
    rw = imp("pde_control_gym.src.rewards")
    ns.NormReward, ns.TunedReward1D = rw.NormReward, rw.TunedReward1D
    ns.NSReward, ns.TrafficARZReward = rw.NSReward, rw.TrafficARZReward
    ns.BrainTumorReward = imp("pde_control_gym.src.rewards.brain_tumor_reward").BrainTumorReward
    ns.norm_reward_module = imp("pde_control_gym.src.rewards.norm_reward")
    _loaded = ns
    return ns


def make_shimmed_norm_reward(ref, nt, norm=2, horizon="temporal", truncate_penalty=-1e-4,
                             terminate_reward=1e2, t_horizon_length=5, norm_coeff=1.0):
    """Reference `NormReward` made runnable WITHOUT editing it.

    The shipped class fails four ways (`norm_reward.py:27,41,54,65`).  The shims:
      * module global `norm_coeff` is injected (undefined name at :54/:59/:61/:73);
      * `uVec` is wrapped in an ndarray subclass whose `== None` is False (:41);
      * `norm` is passed as a number / np.inf rather than the documented string (:54);
      * `t_horizon_length` attribute is aliased (ctor stores the typo `t_hoizon_length`, :27).
    """
    import numpy as np

    ref.norm_reward_module.norm_coeff = norm_coeff

    class _U(np.ndarray):
        def __eq__(self, other):
            if other is None:
                return False
            return np.ndarray.__eq__(self, other)

        __hash__ = None

    inner = ref.NormReward(nt, norm, horizon, truncate_penalty, terminate_reward, t_horizon_length)
    inner.t_horizon_length = t_horizon_length

    class _Shim:
        nt = inner.nt

        def reward(self, uVec=None, time_index=None, terminate=None, truncate=None, action=None):
            return inner.reward(np.asarray(uVec).view(_U), time_index, terminate, truncate, action)

        def reset(self):
            pass

    return _Shim()

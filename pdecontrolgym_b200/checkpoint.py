"""`state_dict()` / `load_state_dict()` for the batched envs: the complete per-env simulation state is a handful of
device tensors (SURVEY.md 5.4 -- the reference envs have no checkpoint support at all)."""
from __future__ import annotations

import torch


class StateDictMixin:
    _STATE_TENSORS: tuple = ()

    def state_dict(self):
        """Copies of every tensor that defines the env batch's state (on the env's device)."""
        out = {}
        for name in self._STATE_TENSORS:
            t = getattr(self, name, None)
            if isinstance(t, torch.Tensor):
                out[name] = t.clone()
        return out

    def load_state_dict(self, sd):
        """Restores a `state_dict()` in place (device addresses are unchanged, so CUDA graphs stay valid)."""
        for name in self._STATE_TENSORS:
            t = getattr(self, name, None)
            if name not in sd:
                continue
            if isinstance(t, torch.Tensor) and t.shape == sd[name].shape and t.dtype == sd[name].dtype:
                t.copy_(sd[name])
            elif t is None or isinstance(t, torch.Tensor):
                raise ValueError(f"state tensor {name!r} has a different shape/dtype than the checkpoint; reset() the env "
                                 "with matching arguments before load_state_dict()")

"""Thin host-side adapters that present a batched env through the two vector-env protocols RL libraries use
(SURVEY.md 8(f) row 1).  Duck-typed on purpose: neither gymnasium nor stable-baselines3 is a dependency.

* `GymnasiumVectorAdapter`: `reset(seed, options) -> (obs, info)`, `step(actions) -> (obs, rewards, terminations,
  truncations, infos)` with numpy arrays (gymnasium.vector.VectorEnv, autoreset mode SAME_STEP).
* `SB3VecEnvAdapter`: `reset() -> obs`, `step_async(actions)` / `step_wait() -> (obs, rewards, dones, infos)` with
  `infos[i]["TimeLimit.truncated"]` (stable_baselines3.common.vec_env.VecEnv).  Terminal observations are not
  retained (same-step auto-reset overwrites the state), so `terminal_observation` is omitted.
"""
from __future__ import annotations

import numpy as np
import torch


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class GymnasiumVectorAdapter:
    def __init__(self, env):
        self.env = env
        self.num_envs = env.num_envs
        self.single_observation_space = getattr(env, "single_observation_space", None)
        self.single_action_space = getattr(env, "single_action_space", None)
        self.observation_space = getattr(env, "observation_space", None)
        self.action_space = getattr(env, "action_space", None)

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return _np(obs), info

    def step(self, actions):
        obs, rew, term, trunc, info = self.env.step(actions)
        return _np(obs), _np(rew), _np(term), _np(trunc), {k: _np(v) if isinstance(v, torch.Tensor) else v for k, v in (info or {}).items()}

    def close(self):
        self.env.close()


class SB3VecEnvAdapter:
    def __init__(self, env):
        if not getattr(env, "autoreset", False):
            raise ValueError("SB3 vector envs auto-reset: construct the env with autoreset=True")
        self.env = env
        self.num_envs = env.num_envs
        self.observation_space = getattr(env, "single_observation_space", None)
        self.action_space = getattr(env, "single_action_space", None)
        self._actions = None

    def reset(self):
        obs, _ = self.env.reset()
        return _np(obs)

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self):
        obs, rew, term, trunc, _ = self.env.step(self._actions)
        term, trunc = _np(term), _np(trunc)
        dones = term | trunc
        infos = [{"TimeLimit.truncated": bool(trunc[i] and not term[i])} for i in range(self.num_envs)]
        return _np(obs), _np(rew).astype(np.float32), dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * self.num_envs

    def seed(self, seed=None):
        return [seed] * self.num_envs


class TrajectoryRecorder:
    """Keeps the observations of the last `max_steps` steps on the device -- the batched stand-in for the
    reference's `env.u[t]` / `env.U[t]` history arrays that its plotting notebooks index
    (base_env_1d.py:41, base_env_2d.py:50).  `history()` returns [steps_recorded, num_envs, ...]."""

    def __init__(self, env, max_steps: int):
        self.env, self.max_steps = env, int(max_steps)
        self._buf, self._n = None, 0

    def reset(self, *args, **kwargs):
        obs, info = self.env.reset(*args, **kwargs)
        self._buf = torch.empty((self.max_steps + 1,) + tuple(obs.shape), dtype=obs.dtype, device=obs.device)
        self._buf[0].copy_(obs)
        self._n = 1
        return obs, info

    def step(self, action):
        out = self.env.step(action)
        if self._n <= self.max_steps:
            self._buf[self._n].copy_(out[0])
            self._n += 1
        return out

    def history(self):
        return self._buf[:self._n]

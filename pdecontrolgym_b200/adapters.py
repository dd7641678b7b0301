"""Thin host-side adapters that present a batched env through the two vector-env protocols RL libraries use
(SURVEY.md 8(f) row 1).  Duck-typed on purpose: neither gymnasium nor stable-baselines3 is a dependency (neither is
installed in the build image; `tests/vecenv_protocol.py` is a vendored checker of the two documented contracts).

* `GymnasiumVectorAdapter`: `reset(seed, options) -> (obs, info)`, `step(actions) -> (obs, rewards, terminations,
  truncations, infos)` with numpy arrays (gymnasium.vector.VectorEnv).  Two auto-reset modes, as in gymnasium >= 1.0:
  - `"same_step"` (needs an env built with `autoreset=True, final_obs=True`): a finished env restarts inside the
    step that finished it; `obs` holds the new episode's first observation and the terminal one is reported as
    `infos["final_obs"][i]` with the mask `infos["_final_obs"]` (gymnasium 0.29 spelling `final_observation` too);
  - `"next_step"` (gymnasium's default; needs `autoreset=False`): the step after the one that finished an env
    ignores that env's action, resets it and returns its first observation with reward 0 and both flags False.
* `SB3VecEnvAdapter`: `reset() -> obs`, `step_async(actions)` / `step_wait() -> (obs, rewards, dones, infos)` with
  `infos[i]["TimeLimit.truncated"]` and, for finished envs, `infos[i]["terminal_observation"]`
  (stable_baselines3.common.vec_env.VecEnv -- what `PPO("MlpPolicy", env)` of the reference's
  examples/transportPDE/transport1Dppo.py:88-90 bootstraps from on truncation).
"""
from __future__ import annotations

import numpy as np
import torch


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _info_np(info):
    return {k: _np(v) if isinstance(v, torch.Tensor) else v for k, v in (info or {}).items()}


class GymnasiumVectorAdapter:
    def __init__(self, env, autoreset_mode: str = "same_step"):
        if autoreset_mode not in ("same_step", "next_step"):
            raise ValueError("autoreset_mode must be 'same_step' or 'next_step'")
        if autoreset_mode == "next_step" and getattr(env, "autoreset", False):
            raise ValueError("next_step auto-reset is driven by the adapter: construct the env with autoreset=False")
        self.env = env
        self.autoreset_mode = autoreset_mode
        self.num_envs = env.num_envs
        self.single_observation_space = getattr(env, "single_observation_space", None)
        self.single_action_space = getattr(env, "single_action_space", None)
        self.observation_space = getattr(env, "observation_space", None)
        self.action_space = getattr(env, "action_space", None)
        self._prev_done = None           # next_step mode: envs that finished in the previous step (device bool tensor)

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        self._prev_done = None
        return _np(obs), _info_np(info)

    def step(self, actions):
        if self.autoreset_mode == "next_step":
            return self._step_next(actions)
        obs, rew, term, trunc, info = self.env.step(actions)
        infos = _info_np(info)
        fin = infos.pop("final_obs", None)
        if fin is not None:
            done = _np(term) | _np(trunc)
            infos["final_obs"] = infos["final_observation"] = fin
            infos["_final_obs"] = infos["_final_observation"] = done
        return _np(obs), _np(rew), _np(term), _np(trunc), infos

    def _step_next(self, actions):
        env, prev = self.env, self._prev_done
        restore = None
        if prev is not None and hasattr(env, "_fenv") and hasattr(env, "simulation_type"):
            restore = env._fenv[:, 0].clone()           # Traffic-ARZ: the ignored step must not advance the float time
        obs, rew, term, trunc, info = env.step(actions)
        if prev is not None:
            # the envs that finished last time are reset instead of stepped: their action is ignored (the batched step
            # leaves a finished env's state alone or the masked reset overwrites it), reward 0, flags False
            if restore is not None:
                env._fenv[:, 0] = torch.where(prev, restore, env._fenv[:, 0])
            robs, _ = env.reset(options={"mask": prev})
            pm = prev.reshape((-1,) + (1,) * (obs.dim() - 1))
            obs = torch.where(pm, robs.to(obs.dtype), obs)
            rew = torch.where(prev, torch.zeros_like(rew), rew)
            term, trunc = term & ~prev, trunc & ~prev
        done = term | trunc
        self._prev_done = done if bool(done.any()) else None
        return _np(obs), _np(rew), _np(term), _np(trunc), _info_np(info)

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
        obs, rew, term, trunc, info = self.env.step(self._actions)
        term, trunc = _np(term), _np(trunc)
        dones = term | trunc
        infos = [{"TimeLimit.truncated": bool(trunc[i] and not term[i])} for i in range(self.num_envs)]
        fin = (info or {}).get("final_obs")
        if fin is not None and dones.any():
            idx = np.nonzero(dones)[0]
            rows = _np(fin[torch.as_tensor(idx, device=fin.device)]) if isinstance(fin, torch.Tensor) else np.asarray(fin)[idx]
            for i, row in zip(idx, rows):          # only the finished envs' rows cross to the host
                infos[i]["terminal_observation"] = row
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
    (base_env_1d.py:41, base_env_2d.py:50).  `history()` returns [steps_recorded, num_envs, ...].

    `extras`: names of per-env attributes of the wrapped env to record alongside, e.g. for BrainTumor1D
    `("dosage", "t1_radius_idx")` -- the reference's `dosage_vs_time` / `t1_radius_idx_vs_time` lists
    (brain_tumor_env.py:54-56,160,276); `extra_history(name)` returns [steps_recorded, num_envs]."""

    def __init__(self, env, max_steps: int, extras=()):
        self.env, self.max_steps = env, int(max_steps)
        self.extras = tuple(extras)
        self._buf, self._n, self._ex = None, 0, {}

    def _grab_extras(self, i):
        for name in self.extras:
            v = getattr(self.env, name)
            v = v() if callable(v) else v
            if name not in self._ex:
                self._ex[name] = torch.empty((self.max_steps + 1,) + tuple(v.shape), dtype=v.dtype, device=v.device)
            self._ex[name][i].copy_(v)

    def reset(self, *args, **kwargs):
        obs, info = self.env.reset(*args, **kwargs)
        self._buf = torch.empty((self.max_steps + 1,) + tuple(obs.shape), dtype=obs.dtype, device=obs.device)
        self._buf[0].copy_(obs)
        self._ex = {}
        self._grab_extras(0)
        self._n = 1
        return obs, info

    def step(self, action):
        out = self.env.step(action)
        if self._n <= self.max_steps:
            self._buf[self._n].copy_(out[0])
            self._grab_extras(self._n)
            self._n += 1
        return out

    def history(self):
        return self._buf[:self._n]

    def extra_history(self, name):
        return self._ex[name][:self._n]

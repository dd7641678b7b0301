"""Vendored checkers of the two vector-env contracts the adapters implement.  gymnasium and stable-baselines3 are not
installed in this image (and must not be installed from an index), so instead of importing their base classes these
functions assert the documented behaviour a `gymnasium.vector.VectorEnv` (v1.x, AutoresetMode.SAME_STEP / NEXT_STEP)
and a `stable_baselines3.common.vec_env.VecEnv` promise to their callers -- shapes, dtypes, info keys and the
auto-reset bookkeeping -- by driving an adapter for a number of steps.  Used on a fake env (CPU tests) and on the CUDA
envs (GPU tests)."""
import numpy as np


def drive_gymnasium(adapter, action_fn, steps, obs_shape, reset_kwargs=None):
    """Steps `adapter` and checks the gymnasium.vector.VectorEnv contract.  Returns the list of step tuples."""
    n = adapter.num_envs
    obs, info = adapter.reset(**(reset_kwargs or {}))
    assert isinstance(obs, np.ndarray) and obs.shape == (n,) + tuple(obs_shape) and isinstance(info, dict)
    first_obs = obs.copy()
    prev_done = np.zeros(n, dtype=bool)
    log = []
    for k in range(steps):
        out = adapter.step(action_fn(k))
        assert isinstance(out, tuple) and len(out) == 5
        obs, rew, term, trunc, infos = out
        assert isinstance(obs, np.ndarray) and obs.shape == (n,) + tuple(obs_shape)
        assert isinstance(rew, np.ndarray) and rew.shape == (n,) and rew.dtype.kind == "f"
        for flag in (term, trunc):
            assert isinstance(flag, np.ndarray) and flag.shape == (n,) and flag.dtype == bool
        assert isinstance(infos, dict)
        done = term | trunc
        if adapter.autoreset_mode == "same_step":
            if done.any():
                # vectorised info convention: infos[key] has one entry per env, infos["_" + key] says which are set
                assert "final_obs" in infos and "_final_obs" in infos
                assert np.array_equal(np.asarray(infos["_final_obs"], dtype=bool), done)
                assert np.asarray(infos["final_obs"]).shape[0] == n
                assert np.array_equal(np.asarray(infos["final_observation"]), np.asarray(infos["final_obs"]))   # 0.29 spelling
        else:
            # next_step: the step after an env finished ignores its action and returns its first observation,
            # reward 0, both flags False
            if prev_done.any():
                assert not term[prev_done].any() and not trunc[prev_done].any()
                assert (rew[prev_done] == 0).all()
                assert np.array_equal(obs[prev_done], first_obs[prev_done])
        prev_done = done
        log.append(out)
    return log


def drive_sb3(adapter, action_fn, steps, obs_shape):
    """Steps `adapter` and checks the stable_baselines3 VecEnv contract.  Returns the list of step tuples."""
    n = adapter.num_envs
    obs = adapter.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (n,) + tuple(obs_shape)
    log = []
    for k in range(steps):
        adapter.step_async(action_fn(k))
        out = adapter.step_wait()
        assert isinstance(out, tuple) and len(out) == 4
        obs, rew, dones, infos = out
        assert isinstance(obs, np.ndarray) and obs.shape == (n,) + tuple(obs_shape)
        assert isinstance(rew, np.ndarray) and rew.shape == (n,) and rew.dtype == np.float32
        assert isinstance(dones, np.ndarray) and dones.shape == (n,) and dones.dtype == bool
        assert isinstance(infos, list) and len(infos) == n and all(isinstance(i, dict) for i in infos)
        for i in range(n):
            assert isinstance(infos[i]["TimeLimit.truncated"], bool)
            if dones[i]:
                # SB3's algorithms bootstrap from this on time-limit truncation (on_policy_algorithm.collect_rollouts)
                assert "terminal_observation" in infos[i]
                assert np.asarray(infos[i]["terminal_observation"]).shape == tuple(obs_shape)
            else:
                assert "terminal_observation" not in infos[i] and not infos[i]["TimeLimit.truncated"]
        log.append(out)
    assert adapter.env_is_wrapped(object) == [False] * n
    return log

"""BrainTumor1D and TherapyWrapper as batched environments (brain_tumor_env.py:33-505 of the reference).

Same constructor keywords as the reference plus `num_envs`, `device`.  `D` and `rho` may be arrays of
shape (num_envs,) (per-patient parameters).  All arithmetic is float64 and runs in
`libpdecg.so` (csrc/brain_tumor.cu); the TherapyWrapper loops (whole Growth stage inside `reset`,
whole Post-Therapy stage inside one `step`, forced weekends, `benchmark`) run inside ONE kernel
launch per call.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from .rewards import BrainTumorReward
from .spaces import Box
from .checkpoint import StateDictMixin

STAGE_NAMES = ("Growth", "Therapy", "Post-Therapy")
RAW_STEP, WRAP_RESET, WRAP_STEP, RUN_TO_END = 0, 1, 2, 3
(I_T, I_STAGE, I_GROWTH, I_THERAPY, I_POST, I_SIM, I_FIRST_THERAPY, I_FIRST_POST, I_DEATH, I_TBENCH,
 I_CONSEC, I_CALLS, I_VIOL) = range(13)
NINT, NF64 = 16, 4
F_REMAINING, F_D, F_RHO, F_LAST_DOSE = range(4)


class BrainTumor1D(StateDictMixin):
    _STATE_TENSORS = ('u', '_obs', '_ienv', '_fenv', '_ic', '_reward', '_terminated', '_truncated', '_valid')
    """Batched `BrainTumor1D`: diffusion-proliferation-radiation PDE with the Growth -> Therapy ->
    Post-Therapy stage machine.  `step(action)` advances every env by one day."""

    def __init__(self,
                 t1_detection_threshold: float = 0.8,
                 t2_detection_threshold: float = 0.16,
                 dosage_termination_threshold: float = 0.1,
                 D=0.2, rho=0.03, alpha: float = 0.04, alpha_beta_ratio: int = 10, k: float = 1e5,
                 t1_detection_radius: int = 15, t1_death_radius: int = 35,
                 reset_init_condition_func=None, total_dosage=None, verbose=False,
                 *, T, dt, X, dx, reward_class=None, normalize=False,
                 num_envs: int = 1, device="cuda", precision: str = "ref", copy_obs: bool = True, autoreset: bool = False,
                 final_obs: bool = False):
        if precision not in ("ref", "fp64", "fp32"):
            raise ValueError("precision must be 'ref' / 'fp64' (float64 like the reference) or 'fp32' (opt-in fast mode: float32 "
                             "state, observation and finite-difference arithmetic)")
        self.precision = precision
        self.state_dtype = torch.float32 if precision == "fp32" else torch.float64
        if reward_class is not None and not isinstance(reward_class, BrainTumorReward):
            raise TypeError("reward_class must be a pdecontrolgym_b200.BrainTumorReward (evaluated inside the kernel)")
        self.num_envs = N = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pdecontrolgym_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.T, self.dt, self.X, self.dx = T, dt, X, dx
        self.nt = int(round(T / dt) + 1)                 # base_env_1d.py:23
        self.nx = int(round(X / dx) + 1)                 # brain_tumor_env.py:52
        self.reward_class = reward_class
        self.verbose = verbose
        self.copy_obs = copy_obs
        self.autoreset = bool(autoreset)
        self.reset_init_condition_func = reset_init_condition_func
        self.total_dosage = float(total_dosage)          # :89 (TypeError on None, like the reference)
        self.k = k
        self.single_action_space = Box(np.full(1, 0, dtype="float32"), np.full(1, 1, dtype="float32"))
        self.action_space = Box(np.full((N, 1), 0, dtype="float32"), np.full((N, 1), 1, dtype="float32"))
        self.single_observation_space = Box(np.full(self.nx, 0, dtype="float64"), np.full(self.nx, k, dtype="float64"), dtype=np.float64)
        self.observation_space = Box(np.full((N, self.nx), 0, dtype="float64"), np.full((N, self.nx), k, dtype="float64"), dtype=np.float64)

        cfg = _lib.BtCfg()
        cfg.abi_version = _lib.ABI_VERSION
        cfg.nx, cfg.nt, cfg.weekends = self.nx, self.nt, 0
        cfg.dt, cfg.dx, cfg.k = float(dt), float(dx), float(k)
        cfg.t1_threshold_k = t1_detection_threshold * k          # :114 detectionRatio * self.k
        cfg.t2_threshold_k = t2_detection_threshold * k
        cfg.dosage_termination_threshold = float(dosage_termination_threshold)
        cfg.alpha, cfg.alpha_beta_ratio = float(alpha), float(alpha_beta_ratio)
        cfg.t1_detection_radius, cfg.t1_death_radius = float(t1_detection_radius), float(t1_death_radius)
        cfg.total_dosage = self.total_dosage
        cfg.precision = _lib.PREC[precision]
        self._cfg = cfg
        self._L = _lib.lib()
        dev = self.device
        with torch.cuda.device(dev):
            self.u = torch.zeros((N, self.nx), dtype=self.state_dtype, device=dev)
            self._obs = torch.zeros((N, self.nx), dtype=self.state_dtype, device=dev)
            self.xScale = torch.from_numpy(np.linspace(0, X, self.nx)).to(dev)      # :57
            self._ienv = torch.zeros((N, NINT), dtype=torch.int32, device=dev)
            self._ienv[:, [I_FIRST_THERAPY, I_FIRST_POST, I_DEATH, I_TBENCH]] = -1
            self._fenv = torch.zeros((N, NF64), dtype=torch.float64, device=dev)
            self._fenv[:, F_REMAINING] = self.total_dosage
            self._fenv[:, F_D] = torch.as_tensor(np.broadcast_to(np.asarray(D, dtype=np.float64), (N,)).copy(), device=dev)
            self._fenv[:, F_RHO] = torch.as_tensor(np.broadcast_to(np.asarray(rho, dtype=np.float64), (N,)).copy(), device=dev)
            self._reward = torch.zeros((N,), dtype=torch.float64, device=dev)
            self._terminated = torch.zeros((N,), dtype=torch.uint8, device=dev)
            self._truncated = torch.zeros((N,), dtype=torch.uint8, device=dev)
            self._valid = torch.ones((N,), dtype=torch.uint8, device=dev)
            self._zero_action = torch.zeros((N,), dtype=torch.float64, device=dev)
            # autoreset + final_obs: the state row of the day that ended an episode, kept for the envs that restarted
            self._final_obs = torch.zeros((N, self.nx), dtype=self.state_dtype, device=dev) if final_obs else None
        buf = _lib.BtBuffers()
        buf.n_envs = N
        buf.u, buf.obs, buf.x_scale = self.u.data_ptr(), self._obs.data_ptr(), self.xScale.data_ptr()
        buf.action = self._zero_action.data_ptr()
        buf.ienv, buf.fenv = self._ienv.data_ptr(), self._fenv.data_ptr()
        buf.reward = self._reward.data_ptr()
        buf.terminated, buf.truncated, buf.valid = self._terminated.data_ptr(), self._truncated.data_ptr(), self._valid.data_ptr()
        self._buf = buf
        self._ic = None

    # -- plumbing ------------------------------------------------------------------------------
    def _stream(self):
        return _lib.current_raw_stream(self.device.index)

    def _run(self, mode, action=None, autoreset=False, mask_t=None, obs_out=None):
        """One launch.  The kernel writes reward / flags / valid straight into fresh tensors handed to the caller (no
        clone launches); returns (reward, terminated_u8, truncated_u8, valid_u8).  `autoreset` (RAW_STEP only): envs
        that finish restart from the last initial condition inside the same launch."""
        a = self._zero_action if action is None else action
        buf = self._buf
        buf.action = a.data_ptr()
        # the raw step's observation is the state itself: a second [N, nx] row is written only when the caller wants a
        # private copy (`obs_out`, a fresh tensor: one launch instead of step + clone)
        buf.obs = (None if obs_out is None else obs_out.data_ptr()) if mode == RAW_STEP else self._obs.data_ptr()
        outs = tuple(torch.empty_like(t) for t in (self._reward, self._terminated, self._truncated, self._valid))
        buf.reward, buf.terminated, buf.truncated, buf.valid = (t.data_ptr() for t in outs)
        # info["stage"] as a private copy: written by the launch itself instead of a clone kernel after it
        self._stage_out = torch.empty((self.num_envs,), dtype=torch.int32, device=self.device) if self.copy_obs else None
        buf.stage_out = None if self._stage_out is None else self._stage_out.data_ptr()
        buf.final_obs = self._final_obs.data_ptr() if (autoreset and self._final_obs is not None) else None
        buf.mask = None if mask_t is None else mask_t.data_ptr()
        try:
            with _lib.device_guard(self.device):
                if autoreset:
                    _lib.check(self._L.pdecg_bt_step_autoreset(C.byref(self._cfg), C.byref(buf), self._ic.data_ptr(),
                                                               0 if self._ic.shape[0] == 1 else self.nx, self._stream()))
                else:
                    _lib.check(self._L.pdecg_bt_run(C.byref(self._cfg), C.byref(buf), mode, self._stream()))
        finally:
            buf.reward, buf.terminated, buf.truncated, buf.valid = (
                self._reward.data_ptr(), self._terminated.data_ptr(), self._truncated.data_ptr(), self._valid.data_ptr())
            buf.final_obs = buf.mask = buf.stage_out = None
        return outs

    def _prep_action(self, action):
        if isinstance(action, torch.Tensor):
            if (action.dtype == torch.float64 and action.dim() == 1 and action.shape[0] == self.num_envs
                    and action.device == self.device and action.is_contiguous()):
                return action                  # already in the kernel's layout: nothing to do
            a = action.to(self.device, torch.float64)
        else:
            a = torch.from_numpy(np.ascontiguousarray(np.asarray(action, dtype=np.float64))).to(self.device)
        a = a.reshape(-1)
        if a.numel() == 1 and self.num_envs > 1:
            a = a.expand(self.num_envs)
        if a.numel() != self.num_envs:
            raise ValueError(f"action must have {self.num_envs} elements (got {a.numel()})")
        return a.contiguous()

    def _outputs(self, obs_src, outs):
        return ((obs_src.clone() if self.copy_obs else obs_src),) + self._out_rest(outs)

    def _out_rest(self, outs):
        """(reward, terminated, truncated, info) from the fresh tensors a launch wrote."""
        reward, term_u8, trunc_u8, valid_u8 = outs
        # copy_obs=False: a live zero-copy view, like the observation; else the private copy the launch wrote
        stage = self._stage_out if self.copy_obs else self._ienv[:, I_STAGE]
        info = {"stage": stage, "valid": valid_u8.view(torch.bool)}
        if self.autoreset and self._final_obs is not None:
            info["final_obs"] = self._final_obs      # rows of the envs with terminated | truncated this step are fresh
        return (reward, term_u8.view(torch.bool), trunc_u8.view(torch.bool), info)

    # -- gym API -------------------------------------------------------------------------------
    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        """brain_tumor_env.py:347-384.  options["init_condition"]: (nx,) or (num_envs, nx) replaces the
        `reset_init_condition_func(X, nx)` callback; options["mask"]: reset only those envs."""
        options = options or {}
        ic = options.get("init_condition")
        if ic is None and self.reset_init_condition_func is None and self._ic is not None:
            ic = self._ic                      # re-use the last explicit initial condition (e.g. benchmark())
        if ic is None:
            if self.reset_init_condition_func is None:
                raise Exception("Please pass an initial condition function")
            try:
                ic = self.reset_init_condition_func(self.X, self.nx, self.num_envs)
            except TypeError:
                rows = [np.asarray(self.reset_init_condition_func(self.X, self.nx)) for _ in range(self.num_envs)]
                ic = rows[0] if all(np.array_equal(r, rows[0]) for r in rows[1:]) else np.stack(rows)
        t = ic.to(self.device, torch.float64) if isinstance(ic, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(ic, dtype=np.float64))).to(self.device)
        if t.dim() == 1:
            t = t.unsqueeze(0)
        if t.shape[-1] != self.nx or t.shape[0] not in (1, self.num_envs):
            raise ValueError(f"init_condition must have shape ({self.nx},) or ({self.num_envs}, {self.nx})")
        t = t.contiguous()
        if self._ic is not None and self._ic is not t and self._ic.shape == t.shape:
            self._ic.copy_(t)          # stable device address across resets (CUDA-graph friendly)
        else:
            self._ic = t
        mask = options.get("mask")
        mask_t = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        with _lib.device_guard(self.device):
            _lib.check(self._L.pdecg_bt_reset(C.byref(self._cfg), C.byref(self._buf), self._ic.data_ptr(),
                                              0 if self._ic.shape[0] == 1 else self.nx,
                                              None if mask_t is None else mask_t.data_ptr(), self._stream()))
        return (self.u.clone() if self.copy_obs else self.u), {}

    def step(self, control):
        """One day for every env (brain_tumor_env.py:123-218).  Envs already at t == nt-1 (where the
        reference returns None) are left unchanged and flagged in info["valid"]."""
        if self._ic is None:
            raise RuntimeError("call reset() before step()")
        # same-step auto-reset is part of the step launch (the returned observation is the restarted state)
        obs = torch.empty_like(self.u) if self.copy_obs else None
        outs = self._run(RAW_STEP, self._prep_action(control), autoreset=self.autoreset, obs_out=obs)
        rest = self._out_rest(outs)
        return ((obs if self.copy_obs else self.u),) + rest

    def step_host(self, action_host, obs_host=None, reward_host=None, terminated_host=None, truncated_host=None, valid_host=None):
        """End-to-end raw step over HOST buffers (pdecg_bt_step_host): H2D(actions) -> kernel -> D2H(state, reward,
        flags) on the current stream, synchronised on return.  CPU tensors (pin them): `action_host` float64 [num_envs];
        `obs_host` float64 [num_envs, nx]; `reward_host` float64 [num_envs]; flags / valid uint8 [num_envs]."""
        if self._ic is None:
            raise RuntimeError("call reset() before step_host()")
        N = self.num_envs
        a_ptr = _lib.host_ptr(action_host, torch.float64, N, "action_host")
        if a_ptr is None:
            raise ValueError("action_host is required")
        if getattr(self, "_act_dev", None) is None:
            self._act_dev = torch.empty((N,), dtype=torch.float64, device=self.device)
        buf = self._buf
        buf.action, buf.obs = self._act_dev.data_ptr(), None
        with _lib.device_guard(self.device):
            _lib.check(self._L.pdecg_bt_step_host(
                C.byref(self._cfg), C.byref(buf), a_ptr,
                _lib.host_ptr(obs_host, self.state_dtype, N * self.nx, "obs_host"),
                _lib.host_ptr(reward_host, torch.float64, N, "reward_host"),
                _lib.host_ptr(terminated_host, torch.uint8, N, "terminated_host"),
                _lib.host_ptr(truncated_host, torch.uint8, N, "truncated_host"),
                _lib.host_ptr(valid_host, torch.uint8, N, "valid_host"),
                self._ic.data_ptr() if self.autoreset else None, 0 if self._ic.shape[0] == 1 else self.nx, self._stream()))
        return obs_host, reward_host, terminated_host, truncated_host, {}

    # -- reference attributes, as per-env tensors ------------------------------------------------
    def _icol(self, i):
        return self._ienv[:, i]

    time_index = property(lambda s: s._icol(I_T))
    stage = property(lambda s: s._icol(I_STAGE))
    growthDays = property(lambda s: s._icol(I_GROWTH))
    therapyDays = property(lambda s: s._icol(I_THERAPY))
    postTherapyDays = property(lambda s: s._icol(I_POST))
    simulationDays = property(lambda s: s._icol(I_SIM))
    firstTherapyDay = property(lambda s: s._icol(I_FIRST_THERAPY))
    firstPostTherapyDay = property(lambda s: s._icol(I_FIRST_POST))
    cDeathDay = property(lambda s: s._icol(I_DEATH))
    remaining_dosage = property(lambda s: s._fenv[:, F_REMAINING])

    # The reference logs two per-day histories, `dosage_vs_time[t]` and `t1_radius_idx_vs_time[t]`
    # (brain_tumor_env.py:54-56,160,276).  Here they are per-env values of the CURRENT day; record them over an episode
    # with `TrajectoryRecorder(env, max_steps, extras=("dosage", "t1_radius_idx"))`.
    @property
    def dosage(self):
        """Dose applied on the last simulated day (0 outside the Therapy stage): the entry the reference writes to
        `dosage_vs_time[time_index]`."""
        return self._fenv[:, F_LAST_DOSE].clone()

    @property
    def t1_radius_idx(self):
        """`getTumorRadius(t, t1_detection_threshold) / dx` of the current state, NaN where the tumour is below the T1
        threshold everywhere: the entry the reference writes to `t1_radius_idx_vs_time[time_index]` (:104-121,274-276)."""
        hit = self.u >= self._cfg.t1_threshold_k
        idx = torch.where(hit, torch.arange(self.nx, device=self.device), -1).max(dim=1).values
        val = (idx.to(torch.float64) * self.dx) / self.dx
        return torch.where(idx >= 0, val, torch.full_like(val, float("nan")))

    @property
    def t_benchmark(self):
        return self._ienv[:, I_TBENCH]

    @t_benchmark.setter
    def t_benchmark(self, value):
        if value is None:
            self._ienv[:, I_TBENCH] = -1
        else:
            self._ienv[:, I_TBENCH] = torch.as_tensor(value, device=self.device).to(torch.int32)

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class TherapyWrapper:
    """Batched `TherapyWrapper` (brain_tumor_env.py:386-505): exposes only the Therapy stage to the
    agent.  `reset` runs the whole Growth stage, a `step` in Post-Therapy runs to the end of the
    episode, `weekends=True` forces two rest days after five consecutive treatment days."""

    def __init__(self, env: BrainTumor1D, weekends=False, verbose=False):
        self.env = env
        self.verbose = verbose
        self.weekends = weekends
        self.num_envs = env.num_envs
        self.action_space, self.observation_space = env.action_space, env.observation_space
        self.single_action_space, self.single_observation_space = env.single_action_space, env.single_observation_space

    @property
    def unwrapped(self):
        return self.env.unwrapped

    treatment_calls = property(lambda s: s.env._ienv[:, I_CALLS])
    soft_constraint_violations = property(lambda s: s.env._ienv[:, I_VIOL])
    consecutive_treatment_days = property(lambda s: s.env._ienv[:, I_CONSEC])

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        env = self.env
        env.reset(seed=seed, options=options)
        env._cfg.weekends = int(bool(self.weekends))
        # a masked reset restarts only the selected patients: the others keep their therapy state, incl. the
        # consecutive-treatment-day counter the weekend rule depends on (brain_tumor_env.py:408-427 is per env)
        mask = (options or {}).get("mask")
        mask_t = None if mask is None else torch.as_tensor(mask, device=env.device).to(torch.uint8).contiguous()
        env._run(WRAP_RESET, mask_t=mask_t)
        return (env._obs.clone() if env.copy_obs else env._obs), {"stage": env._ienv[:, I_STAGE].clone()}

    def step(self, control):
        env = self.env
        env._cfg.weekends = int(bool(self.weekends))
        outs = env._run(WRAP_STEP, env._prep_action(control))
        return env._outputs(env._obs, outs)

    def benchmark(self):
        """Runs a zero-dose episode per env, stores its length in `env.t_benchmark`, resets.  Returns the
        per-env t_benchmark tensor (the reference returns one int)."""
        env = self.env
        env.reset()
        env._run(RUN_TO_END)
        tb = env._ienv[:, I_TBENCH].clone()
        env.reset()
        return tb

"""TransportPDE1D and ReactionDiffusionPDE1D as batched (vector) environments on one B200.

Same constructor keyword names, defaults and error messages as the reference classes
(hyperbolic.py:25-124, parabolic.py:25-124, base_env_1d.py:20-45 of lukebhan/PDEControlGym), plus
`num_envs`, `device`, `precision`.  `reset` / `step` keep the gymnasium signatures but every return
value has a leading `num_envs` axis and lives on the device:

    obs, info = env.reset(seed=None, options=None)
    obs, reward, terminated, truncated, info = env.step(action)      # action: [num_envs] or [num_envs, 1]

All numerical work happens in `libpdecg.so` (csrc/bc1d.cu); this file only owns the device tensors
and fills the C structs.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np
import torch

from . import _lib
from .rewards import NormReward, TunedReward1D
from .spaces import Box
from .checkpoint import StateDictMixin

_CONTROL = {"Dirchilet": 0, "Neumann": 1}
_SENSE_LOC = {"full": 0, "collocated": 1, "opposite": 2}
_SENSE_TYPE = {"Dirchilet": 0, "Neumann": 1}


class UniformScaleIC:
    """On-device random initial condition `u(x, 0) = c * profile(x)`, c ~ U(lo, hi) per env and per episode
    -- the family every reference example draws from (`np.ones(nx) * np.random.uniform(1, 10)`,
    transport1Dppo.py:36-37).  Pass it as `reset(seed=..., options={"init_condition": UniformScaleIC(1, 10)})`;
    with `autoreset=True` every new episode gets a fresh draw.  c comes from Philox4x32-10 keyed by the seed
    with counter (global env id, episode index), so a batch sharded over several GPUs reproduces the
    single-GPU draws (`env_id_offset`)."""

    def __init__(self, lo: float, hi: float, profile=None):
        self.lo, self.hi, self.profile = float(lo), float(hi), profile


class PDEVecEnv1D(StateDictMixin):
    _STATE_TENSORS = ('u', 'beta', '_ic', 'time_index', '_carry', '_episode', '_reward', '_terminated', '_truncated', '_obs_scalar')
    """Batched counterpart of `PDEEnv1D` (base_env_1d.py:9-68)."""

    ENV_KIND = -1
    NPTS_EXTRA = 0

    def __init__(self, T: float, dt: float, X: float, dx: float, reward_class, normalize: bool = False,
                 *, num_envs: int = 1, device="cuda", precision: str = "ref"):
        if precision not in _lib.PREC:
            raise ValueError(f"precision must be one of {sorted(_lib.PREC)} (got {precision!r})")
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pdecontrolgym_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.precision = precision
        self.nt = int(round(T / dt) + 1)
        self.nx = int(round(X / dx))
        self.T, self.dt, self.X, self.dx = T, dt, X, dx
        self.npts = self.nx + self.NPTS_EXTRA
        self.normalize = bool(normalize)
        self.reward_class = reward_class
        self.state_dtype = torch.float64 if precision == "fp64" else torch.float32
        self.single_action_space = Box(np.full(1, -1, dtype="float32"), np.full(1, 1, dtype="float32"))
        self.action_space = Box(np.full((self.num_envs, 1), -1, dtype="float32"),
                                np.full((self.num_envs, 1), 1, dtype="float32"))


class _BoundaryControlVecEnv(PDEVecEnv1D):
    def __init__(self,
                 sensing_noise_func: Optional[Callable] = None,
                 reset_init_condition_func: Optional[Callable] = None,
                 reset_recirculation_func: Optional[Callable] = None,
                 sensing_loc: str = "full",
                 control_type: str = "Dirchilet",
                 sensing_type: str = "Dirchilet",
                 limit_pde_state_size: bool = False,
                 max_state_value: float = 1e10,
                 max_control_value: float = 20,
                 control_sample_rate: float = 0.1,
                 autoreset: bool = False,
                 copy_obs: bool = True,
                 env_id_offset: int = 0,
                 obs_dtype: Optional[str] = None,
                 final_obs: bool = False,
                 truncate_nonfinite: bool = False,
                 report_nonfinite: bool = False,
                 **kwargs):
        super().__init__(**kwargs)
        self.sensing_noise_func = sensing_noise_func
        self.reset_init_condition_func = reset_init_condition_func
        self.reset_recirculation_func = reset_recirculation_func
        self.sensing_loc, self.control_type, self.sensing_type = sensing_loc, control_type, sensing_type
        self.limit_pde_state_size = limit_pde_state_size
        self.max_state_value = max_state_value
        self.max_control_value = max_control_value
        self.control_sample_rate = control_sample_rate
        self.autoreset = bool(autoreset)
        self.copy_obs = bool(copy_obs)   # False: full-state observations are zero-copy views of the live state
        self.env_id_offset = int(env_id_offset)   # global index of env 0 (multi-GPU sharding)
        # obs_dtype="float32": full-state observations (and final_obs) leave the kernels as float32 -- the dtype of the
        # reference's observation_space (base_env_1d.py:33-35) -- while the resident state keeps its precision.  Halves the
        # observation bytes of an fp64 batch (device copy and, for step_host, the PCIe copy-out).
        if obs_dtype not in (None, "state", "float32"):
            raise ValueError("obs_dtype must be None / 'state' (the state dtype) or 'float32'")
        self.obs_f32 = obs_dtype == "float32"
        if self.obs_f32 and not self.copy_obs and sensing_loc == "full":
            raise ValueError("obs_dtype='float32' needs copy_obs=True (a zero-copy view of the state cannot change dtype)")
        self._want_final_obs = bool(final_obs)
        self._seed, self._random_ic = 0, None
        # same messages as hyperbolic.py:60-62,87-89,122-124 / parabolic.py:85,114
        if sensing_loc not in _SENSE_LOC:
            raise Exception("Invalid sensing_loc parameter. Please use 'full', 'collocated', or 'opposite'. See documentation for details.")
        if control_type not in _CONTROL:
            raise Exception("Invalid control_type parameter. Please use 'Neumann' or 'Dirchilet'. See documentation for details.")
        if sensing_loc == "opposite":
            if sensing_type not in _SENSE_TYPE:
                raise Exception("Invalid sensing_type parameter. Please use 'Neumann' or 'Dirchilet'. See documentation for details.")
            if self.ENV_KIND == 1 and sensing_type == "Dirchilet":
                raise Exception("In the parabolic PDE system, u(0, t)=0 and so Dirchilet sensing at u(0, t) is not viable. See documentation for details.")
        N, P, dev = self.num_envs, self.npts, self.device
        obs_n = P if sensing_loc == "full" else 1
        self.single_observation_space = Box(np.full(obs_n, -max_state_value, dtype="float32"),
                                            np.full(obs_n, max_state_value, dtype="float32"))
        self.observation_space = Box(np.full((N, obs_n), -max_state_value, dtype="float32"),
                                     np.full((N, obs_n), max_state_value, dtype="float32"))

        rw = reward_class = self.reward_class
        cfg = _lib.Bc1dCfg()
        cfg.abi_version = _lib.ABI_VERSION
        cfg.env = self.ENV_KIND
        cfg.precision = _lib.PREC[self.precision]
        cfg.beta_dtype = _lib.F32
        cfg.action_dtype = _lib.F32
        cfg.nx, cfg.nt = self.nx, self.nt
        cfg.substeps = int(round(control_sample_rate / self.dt))
        cfg.dt, cfg.dx = float(self.dt), float(self.dx)
        cfg.control_type = _CONTROL[control_type]
        cfg.sensing_loc = _SENSE_LOC[sensing_loc]
        cfg.sensing_type = _SENSE_TYPE.get(sensing_type, 0)
        cfg.normalize = int(self.normalize)
        cfg.max_control_value = float(max_control_value)
        cfg.limit_pde_state_size = int(bool(limit_pde_state_size))
        cfg.max_state_value = float(max_state_value)
        # info["nonfinite"] (report_nonfinite=True or truncate_nonfinite=True): per-env blow-up flag (inf / NaN in the state
        # row); truncate_nonfinite=True also ends the episode of such an env (not reference behaviour: there inf / NaN run on)
        cfg.truncate_nonfinite = int(bool(truncate_nonfinite))
        self.report_nonfinite = bool(report_nonfinite or truncate_nonfinite)
        if isinstance(rw, TunedReward1D):
            cfg.reward_kind = 0
            cfg.norm_ord, cfg.horizon, cfg.t_horizon_length, cfg.norm_coeff = 2, 0, 1, 1.0
        elif isinstance(rw, NormReward):
            cfg.reward_kind = 1
            cfg.norm_ord, cfg.horizon = rw._ord, rw._horizon
            cfg.t_horizon_length, cfg.norm_coeff = int(rw.t_horizon_length), float(rw.norm_coeff)
        else:
            raise TypeError("reward_class must be a pdecontrolgym_b200 TunedReward1D or NormReward instance "
                            f"(got {type(reward_class).__name__}); rewards are evaluated inside the CUDA kernel")
        cfg.reward_nt = int(rw.nt)
        cfg.truncate_penalty = float(rw.truncate_penalty)
        cfg.terminate_reward = float(rw.terminate_reward)
        self._cfg = cfg
        self._L = _lib.lib()
        carry_len = self._L.pdecg_bc1d_carry_len(C.byref(cfg))

        with torch.cuda.device(dev):
            self.u = torch.zeros((N, P), dtype=self.state_dtype, device=dev)
            self._obs_scalar = torch.zeros((N,), dtype=self.state_dtype, device=dev)
            self._reward = torch.zeros((N,), dtype=torch.float64, device=dev)
            self._terminated = torch.zeros((N,), dtype=torch.uint8, device=dev)
            self._truncated = torch.zeros((N,), dtype=torch.uint8, device=dev)
            self.time_index = torch.zeros((N,), dtype=torch.int32, device=dev)
            self._carry = torch.zeros((N, carry_len), dtype=torch.float64, device=dev)
            self._done = torch.zeros((N,), dtype=torch.uint8, device=dev)
            self._episode = torch.zeros((N,), dtype=torch.int32, device=dev)
        self.beta = None
        self._ic = None
        self._kernel_events = None
        self._act = {}
        buf = _lib.Bc1dBuffers()
        buf.n_envs = N
        buf.u = self.u.data_ptr()
        buf.obs = self._obs_scalar.data_ptr()
        buf.reward = self._reward.data_ptr()
        buf.terminated = self._terminated.data_ptr()
        buf.truncated = self._truncated.data_ptr()
        buf.time_index = self.time_index.data_ptr()
        buf.carry = self._carry.data_ptr()
        buf.obs_copy_f32 = int(self.obs_f32)
        self._buf = buf
        self.obs_torch_dtype = torch.float32 if self.obs_f32 else self.state_dtype
        # autoreset + final_obs: the observation of the step that ended an episode survives the fused restart here
        self._final_obs = None
        if self._want_final_obs:
            with torch.cuda.device(dev):
                self._final_obs = torch.zeros((N, P) if sensing_loc == "full" else (N,), dtype=self.obs_torch_dtype, device=dev)
        self._host_obs_dev = None

    # ------------------------------------------------------------------------------------------
    def _stream(self):
        return _lib.current_raw_stream(self.device.index)

    def _call_callback(self, fn, what):
        """Reference callbacks take `nx` and return one row; a batched callback may instead accept
        `(nx, num_envs)` and return `[num_envs, npts]`.  One-row results from a non-batched callback
        are drawn once PER ENV, like `num_envs` independent reference envs would."""
        if fn is None:
            raise Exception("Please pass both an initial condition and a recirculation function in the parameters dictionary. See documentation for more details")
        try:
            out = fn(self.nx, self.num_envs)
        except TypeError:
            rows = [np.asarray(fn(self.nx)) for _ in range(self.num_envs)]
            if all(r is rows[0] or np.array_equal(r, rows[0]) for r in rows[1:]):
                return rows[0]
            return np.stack(rows)
        return out

    def _to_rows(self, x, what):
        """-> (tensor on device [R, npts] with R in {1, num_envs}, env stride in elements)."""
        if isinstance(x, torch.Tensor):
            t = x.to(self.device)
            if t.dtype not in (torch.float32, torch.float64):
                t = t.to(torch.float64)
        else:
            a = np.asarray(x)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)
            t = torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        if t.dim() == 1:
            t = t.unsqueeze(0)
        if t.dim() != 2 or t.shape[1] != self.npts or t.shape[0] not in (1, self.num_envs):
            raise ValueError(f"{what} must have shape ({self.npts},) or ({self.num_envs}, {self.npts}); got {tuple(t.shape)}")
        t = t.contiguous()
        return t, (0 if t.shape[0] == 1 else self.npts)

    def _sense(self):
        obs = self.u if self.sensing_loc == "full" else self._obs_scalar
        if self.obs_f32 and obs.dtype != torch.float32:
            obs = obs.to(torch.float32)
        elif self.copy_obs:
            obs = obs.clone()
        if self.sensing_noise_func is not None:
            obs = self.sensing_noise_func(obs)
        return obs

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        """Batched `reset` (hyperbolic.py:198-227 / parabolic.py:192-221).

        options (all optional): "init_condition" and "beta" -- arrays/tensors of shape (npts,) or
        (num_envs, npts) that replace the constructor callbacks; "mask" -- bool [num_envs], reset only
        those envs."""
        options = options or {}
        ic = options.get("init_condition")
        beta = options.get("beta")
        if ic is None and self.reset_init_condition_func is None and self._ic is not None:
            ic = self._random_ic if self._random_ic is not None else self._ic     # re-use the last specification
        if beta is None and self.reset_recirculation_func is None and self.beta is not None:
            beta = self.beta
        self._random_ic = None
        if isinstance(ic, UniformScaleIC):
            self._random_ic = ic
            if seed is not None:
                self._seed = int(seed)
                if options.get("mask") is None:
                    self._episode.zero_()
            ic = np.ones(self.npts) if ic.profile is None else ic.profile
        if ic is None:
            ic = self._call_callback(self.reset_init_condition_func, "init_condition")
        if beta is None:
            beta = self._call_callback(self.reset_recirculation_func, "beta")
        ic_t, ic_stride = self._to_rows(ic, "init_condition")
        beta_t, beta_stride = self._to_rows(beta, "beta")
        # beta's dtype selects the arithmetic in "ref" mode (hyperbolic.py:146-155); other modes fix it
        if self.precision == "fp64":
            beta_t = beta_t.to(torch.float64)
        elif self.precision == "fp32":
            beta_t = beta_t.to(torch.float32)
        mask = options.get("mask")
        mask_t = None
        if mask is not None:
            mask_t = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            if self.beta is not None and (beta_t.shape != self.beta.shape or beta_t.dtype != self.beta.dtype):
                raise ValueError("masked reset cannot change beta's shape/dtype")
            if self.beta is not None and beta_stride != 0:
                self.beta[mask_t.bool()] = beta_t[mask_t.bool()]
                beta_t = self.beta
        # keep device addresses stable across resets (a captured CUDA graph has them baked in)
        if self.beta is not None and self.beta is not beta_t and self.beta.shape == beta_t.shape and self.beta.dtype == beta_t.dtype:
            self.beta.copy_(beta_t)
            beta_t = self.beta
        if self._ic is not None and self._ic is not ic_t and self._ic.shape == ic_t.shape and self._ic.dtype == ic_t.dtype:
            self._ic.copy_(ic_t)
            ic_t = self._ic
        self.beta = beta_t
        self._ic, self._ic_stride = ic_t, ic_stride
        self._cfg.beta_dtype = _lib.F64 if beta_t.dtype == torch.float64 else _lib.F32
        self._buf.beta = beta_t.data_ptr()
        self._buf.beta_env_stride = beta_stride
        self._reset_kernel(mask_t)
        self._done.zero_()
        return self._sense(), {}

    def _reset_kernel(self, mask_t):
        if self._random_ic is not None:
            r = self._random_ic
            with _lib.device_guard(self.device):
                _lib.check(self._L.pdecg_bc1d_reset_random(
                    C.byref(self._cfg), C.byref(self._buf), self._ic.data_ptr(),
                    _lib.F64 if self._ic.dtype == torch.float64 else _lib.F32, r.lo, r.hi,
                    self._seed & 0xFFFFFFFFFFFFFFFF, self.env_id_offset, self._episode.data_ptr(),
                    None if mask_t is None else mask_t.data_ptr(), self._stream()))
            return
        with _lib.device_guard(self.device):
            _lib.check(self._L.pdecg_bc1d_reset(
                C.byref(self._cfg), C.byref(self._buf), self._ic.data_ptr(),
                _lib.F64 if self._ic.dtype == torch.float64 else _lib.F32, self._ic_stride,
                None if mask_t is None else mask_t.data_ptr(), self._stream()))

    def _autoreset_desc(self):
        key = (self._ic.data_ptr(), self._ic.dtype, self._random_ic, self._seed, self._ic_stride)
        if getattr(self, "_ar_key", None) == key:      # rebuilt only when reset() changed the restart source
            return self._ar_cached
        ar = _lib.Bc1dAutoReset()
        ar.init = self._ic.data_ptr()
        ar.init_dtype = _lib.F64 if self._ic.dtype == torch.float64 else _lib.F32
        r = self._random_ic
        if r is not None:
            ar.random, ar.lo, ar.hi = 1, r.lo, r.hi
            ar.seed, ar.env_id_offset, ar.episode = self._seed & 0xFFFFFFFFFFFFFFFF, self.env_id_offset, self._episode.data_ptr()
        else:
            ar.random, ar.init_env_stride = 0, self._ic_stride
        ar.final_obs = None if self._final_obs is None else self._final_obs.data_ptr()
        self._ar_key, self._ar_cached = key, ar
        return ar

    def _prep_action(self, action):
        if isinstance(action, torch.Tensor):
            a = action
            if (a.dim() == 1 and a.shape[0] == self.num_envs and a.device == self.device and a.is_contiguous()
                    and (a.dtype == torch.float32 or a.dtype == torch.float64)):
                return a                  # the common case: nothing to convert
            if a.dtype not in (torch.float32, torch.float64):
                a = a.to(torch.float32)
            a = a.to(self.device)
        else:
            arr = np.asarray(action)
            if arr.dtype not in (np.float32, np.float64):
                arr = arr.astype(np.float64 if arr.dtype.kind == "f" else np.float32)
            a = torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)
        a = a.reshape(-1)
        if a.numel() == 1 and self.num_envs > 1:
            a = a.expand(self.num_envs)
        if a.numel() != self.num_envs:
            raise ValueError(f"action must have {self.num_envs} elements (got {a.numel()})")
        return a.contiguous()

    def step(self, action):
        """Batched `step` (hyperbolic.py:126-169 / parabolic.py:126-164): `substeps` explicit updates
        with the action held, then terminate / truncate / sensing / reward -- one kernel launch."""
        if self.beta is None:
            raise RuntimeError("call reset() before step()")
        a = self._prep_action(action)
        self._cfg.action_dtype = _lib.F64 if a.dtype == torch.float64 else _lib.F32
        self._buf.action = a.data_ptr()
        # full-state observation copies are written by the kernels themselves into a fresh tensor (no clone launch)
        obs_out = None
        if self.copy_obs and self.sensing_loc == "full":
            obs_out = torch.empty_like(self.u, dtype=self.obs_torch_dtype)
            self._buf.obs_copy = obs_out.data_ptr()
        # the kernel writes reward / flags straight into fresh tensors handed to the caller (no clone launches)
        reward = torch.empty_like(self._reward)
        flags = torch.empty((3 if self.report_nonfinite else 2, self.num_envs), dtype=torch.uint8, device=self.device)   # one allocation
        term_u8, trunc_u8 = flags[0], flags[1]
        self._buf.reward, self._buf.terminated, self._buf.truncated = reward.data_ptr(), term_u8.data_ptr(), trunc_u8.data_ptr()
        self._buf.nonfinite = flags[2].data_ptr() if self.report_nonfinite else None
        ke = self._kernel_events          # optional (start, stop) CUDA events bracketing only the step kernel
        try:
            with _lib.device_guard(self.device):
                if ke is not None:
                    ke[0].record()
                if self.autoreset:          # same-step auto-reset inside the step launch: finished envs restart from their IC
                    _lib.check(self._L.pdecg_bc1d_step_autoreset(C.byref(self._cfg), C.byref(self._buf),
                                                                 C.byref(self._autoreset_desc()), self._stream()))
                else:
                    _lib.check(self._L.pdecg_bc1d_step(C.byref(self._cfg), C.byref(self._buf), self._stream()))
                if ke is not None:
                    ke[1].record()
            terminated, truncated = term_u8.view(torch.bool), trunc_u8.view(torch.bool)
            # rows of the envs with terminated | truncated this step are fresh (SB3 "terminal_observation")
            info = {"final_obs": self._final_obs} if (self.autoreset and self._final_obs is not None) else {}
            if self.report_nonfinite:
                info["nonfinite"] = flags[2].view(torch.bool)
        finally:
            self._buf.obs_copy = None            # never leave a pointer to a caller-owned tensor behind
            self._buf.nonfinite = None
            self._buf.reward, self._buf.terminated, self._buf.truncated = (
                self._reward.data_ptr(), self._terminated.data_ptr(), self._truncated.data_ptr())
        if obs_out is not None:
            obs = obs_out if self.sensing_noise_func is None else self.sensing_noise_func(obs_out)
            return obs, reward, terminated, truncated, info
        return self._sense(), reward, terminated, truncated, info

    def step_host(self, action_host, obs_host=None, reward_host=None, terminated_host=None, truncated_host=None):
        """End-to-end step over HOST buffers (pdecg_bc1d_step_host): H2D(actions) -> kernel ->
        D2H(obs, reward, flags) on the current stream, synchronised on return.  All arguments are CPU
        tensors (pin them for full PCIe bandwidth); outputs are written in place and returned.
        `action_host`: [num_envs] float32/float64.  `obs_host`: [num_envs, npts] (full sensing) or
        [num_envs] in the state dtype."""
        if self.beta is None:
            raise RuntimeError("call reset() before step_host()")
        a = action_host
        if not isinstance(a, torch.Tensor) or a.device.type != "cpu" or not a.is_contiguous():
            raise TypeError("action_host must be a contiguous CPU tensor")
        if a.dtype not in (torch.float32, torch.float64) or a.numel() != self.num_envs:
            raise ValueError("action_host must be float32/float64 with num_envs elements")
        dev_a = self._act.get(a.dtype)
        if dev_a is None:
            dev_a = self._act[a.dtype] = torch.empty(self.num_envs, dtype=a.dtype, device=self.device)
        self._cfg.action_dtype = _lib.F64 if a.dtype == torch.float64 else _lib.F32
        self._buf.action = dev_a.data_ptr()
        def ptr(t, dtype, numel, what):
            if t is None:
                return None
            if t.device.type != "cpu" or not t.is_contiguous() or t.dtype != dtype or t.numel() != numel:
                raise ValueError(f"{what} must be a contiguous CPU {dtype} tensor with {numel} elements")
            return t.data_ptr()
        obs_n = self.num_envs * (self.npts if self.sensing_loc == "full" else 1)
        full = self.sensing_loc == "full"
        host_obs_dtype = self.obs_torch_dtype if full else self.state_dtype
        if full and self.obs_f32 and self.state_dtype != torch.float32 and obs_host is not None:
            # float32 observations of a float64 state: the kernel writes them into this device staging buffer
            if self._host_obs_dev is None:
                self._host_obs_dev = torch.empty((self.num_envs, self.npts), dtype=torch.float32, device=self.device)
            self._buf.obs_copy = self._host_obs_dev.data_ptr()
        try:
            self._step_host_call(a, ptr(obs_host, host_obs_dtype, obs_n, "obs_host"), ptr, reward_host, terminated_host, truncated_host)
        finally:
            self._buf.obs_copy = None
        return obs_host, reward_host, terminated_host, truncated_host, {}

    def _step_host_call(self, a, obs_ptr, ptr, reward_host, terminated_host, truncated_host):
        with _lib.device_guard(self.device):
            _lib.check(self._L.pdecg_bc1d_step_host(
                C.byref(self._cfg), C.byref(self._buf), a.data_ptr(),
                obs_ptr,
                ptr(reward_host, torch.float64, self.num_envs, "reward_host"),
                ptr(terminated_host, torch.uint8, self.num_envs, "terminated_host"),
                ptr(truncated_host, torch.uint8, self.num_envs, "truncated_host"),
                C.byref(self._autoreset_desc()) if self.autoreset else None, self._stream()))

    # compatibility attributes read by the reference's examples
    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class TransportPDE1D(_BoundaryControlVecEnv):
    r"""Batched `TransportPDE1D` (hyperbolic.py): :math:`u_t = u_x + \beta(x) u(0,t)`, first-order
    upwind, boundary control at x = X."""
    ENV_KIND = 0
    NPTS_EXTRA = 0


class ReactionDiffusionPDE1D(_BoundaryControlVecEnv):
    r"""Batched `ReactionDiffusionPDE1D` (parabolic.py): :math:`u_t = u_{xx} + \lambda(x) u`, explicit
    FTCS, u(0,t)=0, ghost point at index nx.  Note the reference's own default
    `control_sample_rate=0.0001` (parabolic.py:34)."""
    ENV_KIND = 1
    NPTS_EXTRA = 1

    def __init__(self, *args, control_sample_rate: float = 0.0001, **kwargs):
        super().__init__(*args, control_sample_rate=control_sample_rate, **kwargs)

"""TrafficPDE1D (Aw-Rascle-Zhang traffic model) as a batched environment
(traffic_arz_env.py:23-280 of the reference).  float64 like the reference (or the opt-in float32 mode, precision='fp32').

The reference's time bookkeeping is kept as is: `time_index` is a float time that
advances by dt per *step*, the PDE integrates only while it is < T, `terminated` fires when it reaches
T/dt (and zeroes it), so with the tutorial defaults an episode has 3,840 steps of which the first 959
integrate.
"""
from __future__ import annotations

import ctypes as C
import random
from typing import Optional

import numpy as np
import torch

from . import _lib
from .rewards import TrafficARZReward
from .spaces import Box
from .checkpoint import StateDictMixin

_TYPES = {"inlet": 0, "outlet": 1, "both": 2, "outlet-train": 3}
F_TIME, F_RS, F_VS, F_QS, F_ACT_LO, F_ACT_HI, NF64 = 0, 1, 2, 3, 4, 5, 6
_RS_VALUES = (0.115, 0.12, 0.125)            # traffic_arz_env.py:66-68


class TrafficPDE1D(StateDictMixin):
    _STATE_TENSORS = ('_ry', '_fenv', '_reward', '_terminated', '_truncated')
    def __init__(self, simulation_type: str = "inlet", v_steady: float = 10, ro_steady: float = 0.12,
                 v_max: float = 40, ro_max: float = 0.16, tau: float = 60,
                 limit_pde_state_size: bool = False, control_freq: int = 1,
                 *, T, dt, X, dx, reward_class=None, normalize=False,
                 num_envs: int = 1, device="cuda", precision: str = "ref", copy_obs: bool = True, autoreset: bool = False,
                 rs_init=None, final_obs: bool = False):
        if precision not in ("ref", "fp64", "fp32"):
            raise ValueError("precision must be 'ref' / 'fp64' (float64 like the reference: bit-identical) or 'fp32' (opt-in fast "
                             "mode: float32 state, observation and arithmetic)")
        self.precision = precision
        self.state_dtype = torch.float32 if precision == "fp32" else torch.float64
        if reward_class is not None and not isinstance(reward_class, TrafficARZReward):
            raise TypeError("reward_class must be a pdecontrolgym_b200.TrafficARZReward (evaluated inside the kernel)")
        assert isinstance(control_freq, int) and control_freq >= 1, \
            f"control_freq must be a positive integer (got {control_freq} of type {type(control_freq).__name__})"
        if simulation_type == "inlet-train":
            raise NotImplementedError("'inlet-train' crashes at the first step in the reference (q_inlet is never set); unsupported")
        if simulation_type not in _TYPES:
            raise ValueError("Invalid simulation type")
        self.num_envs = N = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pdecontrolgym_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.T, self.dt, self.X, self.dx = T, dt, X, dx
        self.nt = int(round(T / dt) + 1)
        self.nx = int(round(X / dx))
        self.simulation_type = simulation_type
        self.vm, self.rm, self.tau = v_max, ro_max, tau
        self.qm = v_max * ro_max / 4
        self.limit_pde_state_size = limit_pde_state_size
        self.control_freq = control_freq
        self.copy_obs = copy_obs
        self.autoreset = bool(autoreset)
        self.reward_class = reward_class
        if simulation_type != "outlet-train":
            if v_steady != self.Veq(v_max, ro_max, ro_steady):
                raise ValueError("The steady state velocity and density do not satisfy the equilibrium condition. Check the values of v_steady and ro_steady and ensure that they obey v_steady = v_max(1 - ro_steady/v_max).")
            rs = np.full(N, float(ro_steady))
            vs = np.full(N, float(v_steady))
            qs = vs * rs                                                   # :63
        else:
            if rs_init is None:           # the constructor's own draw (:66-68); `rs_init` pins it (tests, reproducible runs)
                rs = np.array([_RS_VALUES[random.randint(0, 2)] for _ in range(N)], dtype=np.float64)
            else:
                rs = np.broadcast_to(np.asarray(rs_init, dtype=np.float64), (N,)).copy()
            vs = self.Veq(self.vm, self.rm, rs)
            qs = rs * vs                                                   # :70
        x = np.arange(0, X + dx, dx)                                       # :74
        self.L = X
        self.M = len(x)
        self._profile_np = (np.sin(3 * x / self.L * np.pi) * 0.1 + np.ones([1, self.M])).reshape(-1)   # :83
        nact = 2 if simulation_type == "both" else 1
        cfg = _lib.ArzCfg()
        cfg.abi_version = _lib.ABI_VERSION
        cfg.simulation_type = _TYPES[simulation_type]
        cfg.M, cfg.control_freq = self.M, control_freq
        cfg.limit_pde_state_size = int(bool(limit_pde_state_size))
        cfg.action_dim = nact
        cfg.T, cfg.dt = float(T), float(dt)
        cfg.t_over_dt = T / dt                                             # :109
        cfg.vm, cfg.rm = float(v_max), float(ro_max)
        cfg.c1 = dt / (2 * dx)                                             # :202
        cfg.c2 = 0.25 * dt / tau                                           # :208
        cfg.c3 = dt / dx                                                   # :218
        cfg.c4 = 0.5 * dt / tau                                            # :221
        cfg.vm_over_rm = v_max / ro_max                                    # :84
        cfg.precision = _lib.PREC[precision]
        self._nact = int(cfg.action_dim)
        self._cfg = cfg
        self._L = _lib.lib()
        dev = self.device
        with torch.cuda.device(dev):
            self._ry = torch.zeros((N, 2, self.M), dtype=self.state_dtype, device=dev)
            self._obs = torch.zeros((N, 2 * self.M), dtype=self.state_dtype, device=dev)
            self._fenv = torch.zeros((N, NF64), dtype=torch.float64, device=dev)
            self._profile = torch.from_numpy(self._profile_np.copy()).to(dev)
            self._reward = torch.zeros((N,), dtype=torch.float64, device=dev)
            self._terminated = torch.zeros((N,), dtype=torch.uint8, device=dev)
            self._truncated = torch.zeros((N,), dtype=torch.uint8, device=dev)
            # autoreset + final_obs: the observation of the step that ended an episode, kept for the envs that restarted
            self._final_obs = torch.zeros((N, 2 * self.M), dtype=self.state_dtype, device=dev) if final_obs else None
        self._last_obs = None
        buf = _lib.ArzBuffers()
        buf.n_envs = N
        buf.ry, buf.obs, buf.fenv = self._ry.data_ptr(), self._obs.data_ptr(), self._fenv.data_ptr()
        buf.reward = self._reward.data_ptr()
        buf.terminated, buf.truncated = self._terminated.data_ptr(), self._truncated.data_ptr()
        self._buf = buf
        self._set_steady(rs, vs, qs)
        # action_space.low / .high are fixed here, from the constructor's qs (traffic_arz_env.py:97-100): reset() re-draws
        # rs / qs for 'outlet-train' but the reference keeps clipping with these bounds (:161,166)
        self._fenv[:, F_ACT_LO] = torch.from_numpy(qs * 0.8).to(dev)
        self._fenv[:, F_ACT_HI] = torch.from_numpy(1.2 * qs).to(dev)
        if simulation_type == "outlet-train":
            self.single_observation_space = Box(low=-10, high=10, shape=(2 * self.M,), dtype=np.float64)
        else:
            self.single_observation_space = Box(low=0, high=40, shape=(2 * self.M,), dtype=np.float64)
        self.observation_space = Box(low=self.single_observation_space.low[0], high=self.single_observation_space.high[0],
                                     shape=(N, 2 * self.M), dtype=np.float64)
        q0 = float(qs[0]) if N else float(v_steady * ro_steady)
        self.single_action_space = Box(dtype=np.float64, low=q0 * 0.8, high=1.2 * q0, shape=(nact,))
        self.action_space = Box(dtype=np.float64, low=q0 * 0.8, high=1.2 * q0, shape=(N, nact))
        # the constructor already holds the initial condition; info['V'] is the INITIAL velocity, always (:87-88)
        self._reset_kernel(None)
        self.info = {"V": self._obs[:, self.M:].clone()}

    @staticmethod
    def Veq(vm, rm, rho):
        return vm * (1 - rho / rm)                                         # :270-272

    @staticmethod
    def F_r(vm, rm, rho, y):
        return y + rho * TrafficPDE1D.Veq(vm, rm, rho)

    @staticmethod
    def F_y(vm, rm, rho, y):
        return y * (y / rho + TrafficPDE1D.Veq(vm, rm, rho))

    def _set_steady(self, rs, vs, qs, mask=None):
        vals = torch.from_numpy(np.stack([np.asarray(rs, np.float64), np.asarray(vs, np.float64), np.asarray(qs, np.float64)], axis=1)).to(self.device)
        if mask is None:
            self._fenv[:, F_RS:F_QS + 1] = vals
        else:
            self._fenv[mask, F_RS:F_QS + 1] = vals[mask]

    def _stream(self):
        return _lib.current_raw_stream(self.device.index)

    def _reset_kernel(self, mask_t):
        with _lib.device_guard(self.device):
            _lib.check(self._L.pdecg_arz_reset(C.byref(self._cfg), C.byref(self._buf), self._profile.data_ptr(),
                                               None if mask_t is None else mask_t.data_ptr(), self._stream()))

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        """traffic_arz_env.py:237-267.  options["rs"]: per-env steady densities for 'outlet-train' (default:
        a fresh draw from {0.115, 0.12, 0.125} per env with Python's `random`, like the reference);
        options["mask"]: reset only those envs.  The float time is NOT reset here (reference quirk)."""
        options = options or {}
        mask = options.get("mask")
        mask_t = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        if self.simulation_type == "outlet-train":
            rs = options.get("rs")
            if rs is None:
                # one draw per env that actually resets, in env order (the reference draws once per reset() call, :251)
                which = np.arange(self.num_envs) if mask_t is None else np.flatnonzero(mask_t.cpu().numpy())
                rs = np.full(self.num_envs, _RS_VALUES[1], dtype=np.float64)
                rs[which] = [_RS_VALUES[random.randint(0, 2)] for _ in range(which.size)]
            rs = np.broadcast_to(np.asarray(rs, dtype=np.float64), (self.num_envs,)).copy()
            vs = self.Veq(self.vm, self.rm, rs)
            self._set_steady(rs, vs, rs * vs, None if mask_t is None else mask_t.bool())
        self._reset_kernel(mask_t)
        if not self.copy_obs:
            return self._obs, {}
        if mask_t is not None and self._last_obs is not None:
            # step() hands out fresh observation tensors (self._obs only receives the rows reset() writes), so the
            # rows of the envs that were NOT reset come from the last observation returned
            obs = torch.where(mask_t.bool().unsqueeze(1), self._obs, self._last_obs)
        else:
            obs = self._obs.clone()
        self._last_obs = obs
        return obs, {}

    def step(self, action):
        """traffic_arz_env.py:131-233: clip the flow-rate action(s), `control_freq` Lax-Wendroff substeps,
        reward, terminate / truncate -- one kernel launch for all envs."""
        nact = self._nact
        if (isinstance(action, torch.Tensor) and action.dtype == torch.float64 and action.dim() == 2 and action.shape[0] == self.num_envs
                and action.shape[1] == nact and action.device == self.device and action.is_contiguous()):
            a = action                         # already in the kernel's layout: nothing to do
        else:
            if isinstance(action, torch.Tensor):
                a = action.to(self.device, torch.float64)
            else:
                a = torch.from_numpy(np.ascontiguousarray(np.asarray(action, dtype=np.float64))).to(self.device)
            if a.numel() == nact and self.num_envs != 1:
                a = a.reshape(1, nact).expand(self.num_envs, nact)
            if a.numel() != self.num_envs * nact:
                raise ValueError(f"action must have shape ({self.num_envs}, {nact})")
            a = a.reshape(self.num_envs, nact).contiguous()
        buf = self._buf
        buf.action = a.data_ptr()
        # the kernels write reward / flags (and, with copy_obs, the observation) straight into fresh tensors handed to
        # the caller: no clone launches
        obs = torch.empty_like(self._obs) if self.copy_obs else self._obs
        reward, term_u8, trunc_u8 = torch.empty_like(self._reward), torch.empty_like(self._terminated), torch.empty_like(self._truncated)
        buf.obs, buf.reward, buf.terminated, buf.truncated = obs.data_ptr(), reward.data_ptr(), term_u8.data_ptr(), trunc_u8.data_ptr()
        # same-step auto-reset: finished envs restart from the initial condition inside the step launch (the restart
        # writes the new observation rows where the step wrote).  'outlet-train' re-draws rs per episode on the host, so
        # it keeps the step + masked-reset sequence.
        fused = self.autoreset and self.simulation_type != "outlet-train"
        buf.final_obs = self._final_obs.data_ptr() if (fused and self._final_obs is not None) else None
        info = self.info
        try:
            with _lib.device_guard(self.device):
                if fused:
                    _lib.check(self._L.pdecg_arz_step_autoreset(C.byref(self._cfg), C.byref(buf), self._profile.data_ptr(), self._stream()))
                else:
                    _lib.check(self._L.pdecg_arz_step(C.byref(self._cfg), C.byref(buf), self._stream()))
            terminated, truncated = term_u8.view(torch.bool), trunc_u8.view(torch.bool)
            if self.autoreset and not fused and bool(torch.bitwise_or(term_u8, trunc_u8).any()):
                # ('outlet-train' re-draws the steady state on the host: one sync per step to see whether any env finished,
                # then the masked reset only on the steps that need it)
                done = torch.bitwise_or(term_u8, trunc_u8)
                if self._final_obs is not None:
                    self._final_obs.copy_(torch.where(done.bool().unsqueeze(1), obs, self._final_obs))
                # reward / flags of the restart go to the internal buffers, its observation rows replace the finished ones
                buf.reward, buf.terminated, buf.truncated = self._reward.data_ptr(), self._terminated.data_ptr(), self._truncated.data_ptr()
                buf.obs = self._obs.data_ptr()
                self._last_obs = obs
                obs, _ = self.reset(options={"mask": done})
            if self.autoreset and self._final_obs is not None:
                info = dict(self.info, final_obs=self._final_obs)   # rows of the envs that finished this step are fresh
        finally:
            buf.obs, buf.reward, buf.terminated, buf.truncated = (self._obs.data_ptr(), self._reward.data_ptr(),
                                                                  self._terminated.data_ptr(), self._truncated.data_ptr())
            buf.final_obs = None
        if self.copy_obs:
            self._last_obs = obs
        return obs, reward, terminated, truncated, info

    def step_host(self, action_host, obs_host=None, reward_host=None, terminated_host=None, truncated_host=None):
        """End-to-end step over HOST buffers (pdecg_arz_step_host): H2D(actions) -> kernel -> D2H(obs, reward, flags)
        on the current stream, synchronised on return.  All arguments are CPU tensors (pin them for full PCIe bandwidth):
        `action_host` float64 [num_envs, action_dim]; `obs_host` float64 [num_envs, 2*M]; `reward_host` float64
        [num_envs]; flags uint8 [num_envs].  Outputs are written in place and returned."""
        N, nact = self.num_envs, self._cfg.action_dim
        a_ptr = _lib.host_ptr(action_host, torch.float64, N * nact, "action_host")
        if a_ptr is None:
            raise ValueError("action_host is required")
        if getattr(self, "_act_dev", None) is None:
            self._act_dev = torch.empty((N, nact), dtype=torch.float64, device=self.device)
        buf = self._buf
        buf.action = self._act_dev.data_ptr()
        fused = self.autoreset and self.simulation_type != "outlet-train"
        if self.autoreset and not fused:
            raise NotImplementedError("step_host with autoreset is not available for 'outlet-train' (rs is re-drawn on the host)")
        with _lib.device_guard(self.device):
            _lib.check(self._L.pdecg_arz_step_host(
                C.byref(self._cfg), C.byref(buf), a_ptr,
                _lib.host_ptr(obs_host, self.state_dtype, N * 2 * self.M, "obs_host"),
                _lib.host_ptr(reward_host, torch.float64, N, "reward_host"),
                _lib.host_ptr(terminated_host, torch.uint8, N, "terminated_host"),
                _lib.host_ptr(truncated_host, torch.uint8, N, "truncated_host"),
                self._profile.data_ptr() if fused else None, self._stream()))
        return obs_host, reward_host, terminated_host, truncated_host, self.info

    # reference attributes
    r = property(lambda s: s._ry[:, 0, :])
    y = property(lambda s: s._ry[:, 1, :])
    rs = property(lambda s: s._fenv[:, F_RS])
    vs = property(lambda s: s._fenv[:, F_VS])
    qs = property(lambda s: s._fenv[:, F_QS])
    time_index = property(lambda s: s._fenv[:, F_TIME])

    @property
    def v(self):
        r, y = self.r.double(), self.y.double()
        return y / r + self.vm * (1 - r / self.rm)

    @property
    def ps(self):
        return self.vm / self.rm * self.qs / self.vs                      # :64

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass

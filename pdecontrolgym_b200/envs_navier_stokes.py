"""NavierStokes2D as a batched environment (environments2d/navier_stokes2D.py:9-194 and
base_env_2d.py:24-63 of the reference).  float64 like the reference (or the opt-in float32 mode, precision='fp32'), arrays indexed [env, row(y), col(x)].

One scalar action per env drives every "Controllable" boundary (the reference's `action_dim=1` use).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from .rewards import NSReward
from .spaces import Box
from .checkpoint import StateDictMixin

_BC = {"Dirchilet": 0, "Neumann": 1, "Controllable": 2}
_SIDES = ("lower", "upper", "left", "right")          # the order apply_boundary visits them (:76)


def central_difference(f, coordinate, step=0.01):
    """Batched torch version of the reference helper (navier_stokes2D.py:9-15); f: [..., ny, nx]."""
    diff = torch.zeros_like(f)
    if coordinate == "x":
        diff[..., 1:-1, 1:-1] = (f[..., 1:-1, 2:] - f[..., 1:-1, 0:-2]) / (2 * step)
    elif coordinate == "y":
        diff[..., 1:-1, 1:-1] = (f[..., 2:, 1:-1] - f[..., 0:-2, 1:-1]) / (2 * step)
    return diff


def laplace(f, dx=0.01, dy=0.01):
    """Batched torch version of the reference helper (navier_stokes2D.py:17-22)."""
    diff = torch.zeros_like(f)
    diff[..., 1:-1, 1:-1] = (f[..., 1:-1, 0:-2] + f[..., 0:-2, 1:-1] - 4 * f[..., 1:-1, 1:-1]
                             + f[..., 1:-1, 2:] + f[..., 2:, 1:-1]) / (dx * dy)
    return diff


class NavierStokes2D(StateDictMixin):
    _STATE_TENSORS = ('u', 'v', 'p', '_obs', 'time_index', '_reward', '_terminated', '_truncated', '_valid')
    def __init__(self, reset_init_condition_func=None, boundary_condition: dict = None, U_ref=None, action_ref=None,
                 viscosity: float = 0.1, density: float = 1.0, maximum_pressure_iteration: float = 2000,
                 stable_factor: float = 0.5,
                 *, T, dt, X, dx, Y, dy, action_dim: int = 1, reward_class=None, normalize: bool = False,
                 num_envs: int = 1, device="cuda", precision: str = "ref", copy_obs: bool = True, autoreset: bool = False,
                 final_obs: bool = False):
        if precision not in ("ref", "fp64", "fp32"):
            raise ValueError("NavierStokes2D: precision must be 'ref' / 'fp64' (float64 as the reference, bit-identical fields) "
                             "or 'fp32' (opt-in fast mode: float32 fields and arithmetic)")
        self.precision = precision
        # 'fp32': float32 u / v / p / observation and float32 arithmetic in the same operation order; actions, U_ref,
        # action_ref, the reward and the initial fields stay float64 (tests/test_gpu_navier_stokes.py states the tolerance)
        self.state_dtype = torch.float32 if precision == "fp32" else torch.float64
        if not isinstance(reward_class, NSReward):
            raise TypeError("reward_class must be a pdecontrolgym_b200.NSReward (evaluated inside the kernel)")
        if action_dim != 1:
            raise NotImplementedError("only action_dim=1 (one scalar per env for all Controllable sides) is supported")
        self.num_envs = N = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pdecontrolgym_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.nt = int(round(T / dt))                      # base_env_2d.py:27
        self.nx = int(round(X / dx + 1))
        self.ny = int(round(Y / dy + 1))
        if self.nx != self.ny:
            raise ValueError("the reference's boundary indexing only works on square grids (nx == ny)")
        self.dx, self.dy, self.dt = dx, dy, dt
        self.x = np.linspace(0, X, self.nx)
        self.y = np.linspace(0, Y, self.ny)
        self.X, self.Y = np.meshgrid(self.x, self.y)
        self.reset_init_condition_func = reset_init_condition_func
        self.KINEMATIC_VISCOSITY, self.DENSITY = viscosity, density
        self.N_PRESSURE_POISSON_ITERATIONS = int(maximum_pressure_iteration)
        self.reward_class = reward_class
        self.copy_obs = copy_obs
        self.autoreset = bool(autoreset)
        self._init = None
        self._want_final_obs = bool(final_obs)
        max_t = (0.5 * min(dx, dy) ** 2 / viscosity)
        if dt > stable_factor * max_t:
            raise RuntimeError("Stability is not guarenteed")      # :56-58
        self.boundary_condition = boundary_condition
        cfg = _lib.NsCfg()
        cfg.abi_version = _lib.ABI_VERSION
        cfg.nx, cfg.ny, cfg.nt = self.nx, self.ny, self.nt
        cfg.n_pressure_iters = self.N_PRESSURE_POISSON_ITERATIONS
        for s, side in enumerate(_SIDES):
            for c in range(2):
                cond = boundary_condition[side][c]
                if cond not in _BC:
                    raise ValueError(f"unknown boundary condition {cond!r} (use 'Dirchilet', 'Neumann' or 'Controllable')")
                cfg.bc[s][c] = _BC[cond]
        cfg.dt, cfg.dx, cfg.dy = float(dt), float(dx), float(dy)
        cfg.viscosity, cfg.density, cfg.gamma = float(viscosity), float(density), float(reward_class.gamma)
        cfg.precision = _lib.PREC[precision]
        dev = self.device
        uref = np.ascontiguousarray(np.asarray(U_ref, dtype=np.float64))
        if uref.ndim != 4 or uref.shape[1:] != (self.nx, self.ny, 2):
            raise ValueError(f"U_ref must have shape (nt, {self.nx}, {self.ny}, 2)")
        aref = np.ascontiguousarray(np.asarray(action_ref, dtype=np.float64).reshape(-1))
        cfg.n_uref, cfg.n_action_ref = uref.shape[0], aref.shape[0]
        self._cfg = cfg
        self._L = _lib.lib()
        with torch.cuda.device(dev):
            self.U_ref = torch.from_numpy(uref).to(dev)
            self.action_ref = torch.from_numpy(aref).to(dev)
            shp = (N, self.ny, self.nx)
            self.u = torch.zeros(shp, dtype=self.state_dtype, device=dev)
            self.v = torch.zeros(shp, dtype=self.state_dtype, device=dev)
            self.p = torch.zeros(shp, dtype=self.state_dtype, device=dev)
            self._obs = torch.zeros(shp + (2,), dtype=self.state_dtype, device=dev)
            self.time_index = torch.zeros((N,), dtype=torch.int32, device=dev)
            self._reward = torch.zeros((N,), dtype=torch.float64, device=dev)
            self._terminated = torch.zeros((N,), dtype=torch.uint8, device=dev)
            self._truncated = torch.zeros((N,), dtype=torch.uint8, device=dev)
            self._valid = torch.ones((N,), dtype=torch.uint8, device=dev)
        buf = _lib.NsBuffers()
        buf.n_envs = N
        buf.u, buf.v, buf.p, buf.obs = self.u.data_ptr(), self.v.data_ptr(), self.p.data_ptr(), self._obs.data_ptr()
        buf.u_ref, buf.action_ref = self.U_ref.data_ptr(), self.action_ref.data_ptr()
        buf.time_index, buf.reward = self.time_index.data_ptr(), self._reward.data_ptr()
        buf.terminated, buf.truncated, buf.valid = self._terminated.data_ptr(), self._truncated.data_ptr(), self._valid.data_ptr()
        # grids beyond 64 x 64 do not fit one SM's shared memory: their step streams tiles through HBM and needs scratch
        # (predictor fields, right-hand side, pressure ping-pong), owned here like every other device buffer
        wbytes = int(self._L.pdecg_ns_workspace_bytes(C.byref(cfg), N))
        self._work = torch.empty((wbytes + 7) // 8, dtype=torch.float64, device=dev) if wbytes else None
        buf.work = None if self._work is None else self._work.data_ptr()
        self._buf = buf
        self._reset_done = False
        self.single_observation_space = Box(np.full((self.nx, self.ny, 2), -np.inf, dtype="float32"),
                                            np.full((self.nx, self.ny, 2), np.inf, dtype="float32"))
        self.observation_space = Box(np.full((N, self.nx, self.ny, 2), -np.inf, dtype="float32"),
                                     np.full((N, self.nx, self.ny, 2), np.inf, dtype="float32"))
        self.single_action_space = Box(low=-1.0, high=1.0, shape=(action_dim,), dtype=np.float32)
        self.action_space = Box(low=-1.0, high=1.0, shape=(N, action_dim), dtype=np.float32)

    def _stream(self):
        return _lib.current_raw_stream(self.device.index)

    def _field(self, x, what):
        t = x.to(self.device, torch.float64) if isinstance(x, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(self.device)
        if t.dim() == 2:
            t = t.unsqueeze(0)
        if tuple(t.shape[1:]) != (self.ny, self.nx) or t.shape[0] not in (1, self.num_envs):
            raise ValueError(f"{what} must have shape ({self.ny}, {self.nx}) or ({self.num_envs}, {self.ny}, {self.nx})")
        return t.expand(self.num_envs, self.ny, self.nx).contiguous() if t.shape[0] == 1 and self.num_envs != 1 else t.contiguous()

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        """navier_stokes2D.py:171-194.  options["init"] = (u, v, p) arrays of shape (ny, nx) or
        (num_envs, ny, nx) replaces the `reset_init_condition_func(meshgrid_X)` callback (called once per env
        when it returns single fields)."""
        options = options or {}
        init = options.get("init")
        if init is None and self.reset_init_condition_func is None and self._init is not None:
            init = self._init                  # re-use the last explicit initial fields (masked restarts)
        if init is None:
            if self.reset_init_condition_func is None:
                raise Exception("Please pass both an initial condition and a recirculation function in the parameters dictionary. See documentation for more details")
            trip = [self.reset_init_condition_func(self.X) for _ in range(self.num_envs)]
            init = tuple(np.stack([np.asarray(t[c], dtype=np.float64) for t in trip]) for c in range(3))
        iu, iv, ip = (self._field(f, n) for f, n in zip(init, ("init_u", "init_v", "init_p")))
        if self._init is not None and all(a.shape == b.shape for a, b in zip(self._init, (iu, iv, ip))):
            for dst, src in zip(self._init, (iu, iv, ip)):
                if dst is not src:
                    dst.copy_(src)     # stable device addresses across resets (CUDA-graph friendly)
            iu, iv, ip = self._init
        self._init = (iu, iv, ip)
        mask = options.get("mask")
        mask_t = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        with _lib.device_guard(self.device):
            _lib.check(self._L.pdecg_ns_reset(C.byref(self._cfg), C.byref(self._buf), iu.data_ptr(), iv.data_ptr(), ip.data_ptr(),
                                              self.nx * self.ny, None if mask_t is None else mask_t.data_ptr(), self._stream()))
        self._reset_done = True
        return (self._obs.clone() if self.copy_obs else self._obs), {}

    def step(self, action):
        """navier_stokes2D.py:118-157: predictor, boundary conditions, `maximum_pressure_iteration` Jacobi
        sweeps, corrector, boundary conditions, NSReward -- one kernel launch for all envs."""
        if not self._reset_done:
            raise RuntimeError("call reset() before step()")
        a = action.to(self.device, torch.float64) if isinstance(action, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(action, dtype=np.float64))).to(self.device)
        a = a.reshape(-1)
        if a.numel() == 1 and self.num_envs > 1:
            a = a.expand(self.num_envs)
        if a.numel() != self.num_envs:
            raise ValueError(f"action must have {self.num_envs} elements")
        a = a.contiguous()
        buf = self._buf
        buf.action = a.data_ptr()
        internal = (self._reward, self._terminated, self._truncated, self._valid)
        # the kernel writes reward / flags / valid straight into fresh tensors handed to the caller (no clone launches)
        outs = tuple(torch.empty_like(t) for t in internal)
        buf.reward, buf.terminated, buf.truncated, buf.valid = (t.data_ptr() for t in outs)
        fin = None
        if self.autoreset and self._want_final_obs:
            # the terminal observation U[nt-1] of the envs that restart, kept across the fused auto-reset
            if getattr(self, "_final_obs", None) is None:
                self._final_obs = torch.zeros_like(self._obs)
            fin = self._final_obs
        buf.final_obs = None if fin is None else fin.data_ptr()
        try:
            with _lib.device_guard(self.device):
                if self.autoreset:       # same-step auto-reset to the last initial fields, inside the step launch
                    iu, iv, ip = self._init
                    _lib.check(self._L.pdecg_ns_step_autoreset(C.byref(self._cfg), C.byref(buf), iu.data_ptr(), iv.data_ptr(),
                                                               ip.data_ptr(), self.nx * self.ny, self._stream()))
                else:
                    _lib.check(self._L.pdecg_ns_step(C.byref(self._cfg), C.byref(buf), self._stream()))
        finally:
            buf.reward, buf.terminated, buf.truncated, buf.valid = (t.data_ptr() for t in internal)
            buf.final_obs = None
        reward, term_u8, trunc_u8, valid_u8 = outs
        terminated, truncated, info = term_u8.view(torch.bool), trunc_u8.view(torch.bool), {"valid": valid_u8.view(torch.bool)}
        if fin is not None:
            info["final_obs"] = fin            # rows of the envs with terminated this step are fresh
        obs = self._obs.clone() if self.copy_obs else self._obs
        return obs, reward, terminated, truncated, info

    def step_host(self, action_host, obs_host=None, reward_host=None, terminated_host=None, truncated_host=None, valid_host=None):
        """End-to-end step over HOST buffers (pdecg_ns_step_host): H2D(actions) -> kernel -> D2H(U[t], reward, flags) on
        the current stream, synchronised on return.  CPU tensors (pin them): `action_host` float64 [num_envs];
        `obs_host` [num_envs, ny, nx, 2] in the state dtype (float64; float32 with precision="fp32"); `reward_host` float64
        [num_envs]; flags / valid uint8 [num_envs]."""
        if not self._reset_done:
            raise RuntimeError("call reset() before step_host()")
        N = self.num_envs
        a_ptr = _lib.host_ptr(action_host, torch.float64, N, "action_host")
        if a_ptr is None:
            raise ValueError("action_host is required")
        if getattr(self, "_act_dev", None) is None:
            self._act_dev = torch.empty((N,), dtype=torch.float64, device=self.device)
        buf = self._buf
        buf.action = self._act_dev.data_ptr()
        iu, iv, ip = self._init if self.autoreset else (None, None, None)
        with _lib.device_guard(self.device):
            _lib.check(self._L.pdecg_ns_step_host(
                C.byref(self._cfg), C.byref(buf), a_ptr,
                _lib.host_ptr(obs_host, self.state_dtype, N * self.nx * self.ny * 2, "obs_host"),
                _lib.host_ptr(reward_host, torch.float64, N, "reward_host"),
                _lib.host_ptr(terminated_host, torch.uint8, N, "terminated_host"),
                _lib.host_ptr(truncated_host, torch.uint8, N, "truncated_host"),
                _lib.host_ptr(valid_host, torch.uint8, N, "valid_host"),
                None if iu is None else iu.data_ptr(), None if iv is None else iv.data_ptr(), None if ip is None else ip.data_ptr(),
                self.nx * self.ny, self._stream()))
        return obs_host, reward_host, terminated_host, truncated_host, {}

    def solve_pressure(self, u, v, p_prev):
        """Batched `solve_pressure` (navier_stokes2D.py:94-116) for external callers; returns the new p (float64 only)."""
        if self.precision == "fp32":
            raise NotImplementedError("solve_pressure is float64 only; build the env with precision='ref'")
        u, v = self._field(u, "u"), self._field(v, "v")
        p = self._field(p_prev, "p_prev").clone()
        wbytes = int(self._L.pdecg_ns_workspace_bytes(C.byref(self._cfg), u.shape[0]))
        work = torch.empty((wbytes + 7) // 8, dtype=torch.float64, device=self.device) if wbytes else None
        with _lib.device_guard(self.device):
            _lib.check(self._L.pdecg_ns_solve_pressure(C.byref(self._cfg), u.shape[0], u.data_ptr(), v.data_ptr(), p.data_ptr(),
                                                       None if work is None else work.data_ptr(), self._stream()))
        return p

    @property
    def U(self):
        return self._obs

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass

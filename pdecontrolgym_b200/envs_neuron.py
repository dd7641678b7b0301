"""NeuronPDE1D (neuron growth control) as a batched closed-loop simulation (neuron_env.py:8-372 of the reference).

The reference class is unregistered and cannot run as shipped (`self.atilde1` vs `aTilde1`, no `reset()`, the default
`sensing_loc="Full"` raises); its `step()` takes NO action -- the boundary input is the env's own backstepping law.  This
class keeps the constructor keywords and defaults (with `sensing_loc="full"` as the working default), adds the missing
`reset()`, and advances every env by `n_steps` explicit steps per launch.  Observations are the full error row u(x, t)
[num_envs, M]; the moving boundary index L (points beyond it keep their last values), the ODE state Z and the control input
come back in `info`.  Parity: 1e-10 relative against the numpy oracle (the control law's integral is a warp reduction
tree here, a sequential Python sum in the reference), the oracle is pinned to the shimmed reference at 1e-12."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from .checkpoint import StateDictMixin
from .rewards import NormReward
from .spaces import Box


class NeuronPDE1D(StateDictMixin):
    _STATE_TENSORS = ('u', 'Z', 'L', 'time_index', 'control')

    def __init__(self, sensing_noise_func=None, cInfty: float = 0.0119, LSubS: float = 12e-6, k1: float = -1e3, k2: float = 1e13,
                 LSubZero: float = 1e-6, LSubC: float = 4e-6, a: float = 1e-8, g: float = 5e-7, D: float = 10e-12,
                 TildeRSubG: float = 0.053, RSubG: float = 1.783e-5, limit_pde_state_size: bool = False, max_state_value: float = 1e10,
                 max_control_value: float = 20, control_type: str = "Neumann", sensing_type: str = "Neumann", sensing_loc: str = "full",
                 *, T, dt, X, dx, reward_class=None, normalize: bool = False, num_envs: int = 1, device="cuda", precision: str = "ref"):
        from scipy.linalg import expm
        if precision not in ("ref", "fp64"):
            raise ValueError("NeuronPDE1D is float64 in the reference; precision must be 'ref' or 'fp64'")
        if control_type not in ("Neumann", "Dirchilet"):
            raise Exception("Invalid control_type parameter. Please use 'Neumann' or 'Dirchilet'. See documentation for details.")
        if sensing_loc != "full":
            raise Exception("Invalid sensing_loc parameter. Please use 'full', 'collocated', or 'opposite'. See documentation for details."
                            if sensing_loc not in ("collocated", "opposite") else "only sensing_loc='full' is implemented for NeuronPDE1D")
        if reward_class is not None and not isinstance(reward_class, NormReward):
            raise TypeError("reward_class must be None or a pdecontrolgym_b200.NormReward (reward = -||u[:L]||_p / norm_coeff)")
        self.num_envs = N = int(num_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pdecontrolgym_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.T, self.dt, self.X, self.dx = T, dt, X, dx
        self.nt = int(round(T / dt) + 1)
        self.reward_class, self.sensing_noise_func = reward_class, sensing_noise_func
        self.limit_pde_state_size, self.max_state_value = limit_pde_state_size, max_state_value
        self.cInfty, self.LSubS, self.LSubZero, self.LSubC = cInfty, LSubS, LSubZero, LSubC
        # ---- the constants of neuron_env.py:60-259, formed with the same Python / numpy expressions ----
        gamma = (a / D) * 10
        H = np.zeros([2, 1]); H[0, 0] = 1; H[1, 0] = -1 * (((a - g * LSubC) * (cInfty)) / D)
        aTilde1 = ((a - RSubG * cInfty) / LSubC) - g - TildeRSubG
        Asub1 = np.array([[aTilde1, 0.0], [RSubG, 0.0]])
        beta = D / LSubC
        Bv = np.zeros([2, 1]); Bv[0, 0] = -1 * beta
        s = np.sqrt(a ** 2 + 4 * D * g)
        LambdaMinus, LambdaPlus = (a - s) / (2 * D), (a + s) / (2 * D)
        Kminus, Kplus = 0.5 - ((a - 2 * g * LSubC) / 2) / s, 0.5 + ((a - 2 * g * LSubC) / 2) / s
        K = np.zeros([2, 1]); K[0, 0] = k1; K[1, 0] = k2
        length, scale = X, 1
        while length < 1:
            length, scale = length * 10, scale / 10
        self.length, self.SpatialToRealScale = length, scale
        I2, Z2 = np.eye(2), np.zeros((2, 2))
        N1 = np.block([[Z2, (1.0 / D) * (g * I2 + Asub1 + (a / D) * (Bv @ H.T))], [I2, (1.0 / D) * (Bv @ H.T + a * I2)]])
        Row = np.hstack([H.T, K.T - ((H.T @ Bv @ H.T) / D)])
        Iv, Ivf = np.vstack([I2, np.zeros((2, 2))]), np.vstack([np.zeros((2, 2)), I2])
        x = np.arange(0, length + dx, dx)
        self.M = M = len(x)
        dist = (scale * x) - LSubS
        self.CSubEq = cInfty * (Kplus * np.exp(LambdaPlus * dist) + Kminus * np.exp(LambdaMinus * dist))
        phi = np.stack([(Row @ expm(-1 * scale * (xi) * N1) @ Iv).ravel() for xi in x])
        phip = np.stack([(Row @ expm(-1 * (scale) * (xi) * N1) @ Ivf).ravel() for xi in x])
        p = phip - (gamma * phi)
        pB = np.array([(p[i, :] @ Bv)[0] for i in range(M)])
        cfg = _lib.NeuronCfg()
        cfg.abi_version, cfg.M, cfg.nt = _lib.ABI_VERSION, M, self.nt
        cfg.dt, cfg.dx, cfg.D, cfg.a, cfg.g, cfg.RSubG = float(dt), float(dx), float(D), float(a), float(g), float(RSubG)
        cfg.aTilde1, cfg.beta, cfg.LSubS, cfg.inv_scale = float(aTilde1), float(beta), float(LSubS), float(1 / scale)
        cfg.H0, cfg.H1 = float(H[0, 0]), float(H[1, 0])
        cfg.coef = float(((H.T @ Bv) / D + gamma)[0, 0])
        self._cfg = cfg
        self._L = _lib.lib()
        dev = self.device
        with torch.cuda.device(dev):
            self.u = torch.zeros((N, M), dtype=torch.float64, device=dev)
            self.Z = torch.zeros((N, 2), dtype=torch.float64, device=dev)
            self.L = torch.zeros((N,), dtype=torch.int32, device=dev)
            self.time_index = torch.zeros((N,), dtype=torch.float64, device=dev)
            self.control = torch.zeros((N,), dtype=torch.float64, device=dev)
            self._terminated = torch.zeros((N,), dtype=torch.uint8, device=dev)
            self.p = torch.from_numpy(np.ascontiguousarray(p)).to(dev)
            self._pB = torch.from_numpy(pB).to(dev)
        buf = _lib.NeuronBuffers()
        buf.n_envs = N
        buf.u, buf.Z, buf.L, buf.time = self.u.data_ptr(), self.Z.data_ptr(), self.L.data_ptr(), self.time_index.data_ptr()
        buf.control, buf.pB, buf.p, buf.terminated = self.control.data_ptr(), self._pB.data_ptr(), self.p.data_ptr(), self._terminated.data_ptr()
        self._buf = buf
        self.single_observation_space = Box(np.full(M, -max_state_value, dtype="float32"), np.full(M, max_state_value, dtype="float32"))
        self.observation_space = Box(np.full((N, M), -max_state_value, dtype="float32"), np.full((N, M), max_state_value, dtype="float32"))
        self.single_action_space = Box(low=np.array([-max_control_value], dtype="float32"), high=np.array([max_control_value], dtype="float32"))
        self.action_space = self.single_action_space
        self.reset()

    def _conversion(self, z1):                                                   # neuron_env.py:363-372
        raw = np.rint((float(1 / self.SpatialToRealScale) * (np.asarray(z1) + self.LSubS)) / self.dx).astype(np.int64)
        return np.clip(raw, 1, self.M - 1)

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        """The initial state the reference's constructor builds (:205-225).  options["LSubZero"]: per-env initial axon lengths."""
        N = self.num_envs
        l0 = np.broadcast_to(np.asarray((options or {}).get("LSubZero", self.LSubZero), dtype=np.float64), (N,))
        z = np.stack([np.full(N, self.cInfty), l0 - self.LSubS], axis=1)
        self.u.copy_(torch.from_numpy(np.tile(2 * self.cInfty - self.CSubEq, (N, 1))))
        self.Z.copy_(torch.from_numpy(z))
        self.L.copy_(torch.from_numpy(self._conversion(z[:, 1]).astype(np.int32)))
        self.time_index.zero_()
        self.control.zero_()
        return self.u.clone(), {"L": self.L.clone(), "Z": self.Z.clone()}

    def step(self, action=None, n_steps: int = 1):
        """`n_steps` explicit steps of every env in one launch (neuron_env.py:263-339; `action` is ignored like the reference's
        step() takes none).  Returns (u [N, M], reward [N], terminated [N], truncated [N], info)."""
        with _lib.device_guard(self.device):
            _lib.check(self._L.pdecg_neuron_step(C.byref(self._cfg), C.byref(self._buf), int(n_steps), _lib.current_raw_stream(self.device.index)))
        obs = self.u.clone()
        mask = torch.arange(self.M, device=self.device)[None, :] < self.L[:, None]           # new_u = u[:L]  (:324)
        live = torch.where(mask, obs, torch.zeros_like(obs))
        if self.reward_class is not None:
            reward = -torch.linalg.vector_norm(live, ord={1: 1, 2: 2, 3: float('inf')}[self.reward_class._ord], dim=1) / self.reward_class.norm_coeff
        else:
            reward = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
        truncated = (torch.linalg.vector_norm(obs, dim=1) >= self.max_state_value) if self.limit_pde_state_size else torch.zeros_like(self._terminated, dtype=torch.bool)
        if self.sensing_noise_func is not None:
            obs = self.sensing_noise_func(obs)
        return obs, reward, self._terminated.clone().view(torch.bool), truncated, {"L": self.L.clone(), "Z": self.Z.clone(), "control": self.control.clone()}

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass

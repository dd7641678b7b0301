"""ctypes wrapper of oracle/c/pdecg_oracle.c (plain-C restatement of the Transport / ReactionDiffusion, Traffic-ARZ
and NavierStokes2D integrators) -- TEST INFRASTRUCTURE ONLY; see the C file's header.  State-only (no rewards): it
checks the integrators at full batch sizes and times a compiled multi-threaded CPU port."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "c", "libpdecg_oracle.so")


class Cfg(C.Structure):
    _fields_ = [("env", C.c_int), ("state_f64", C.c_int), ("beta_f64", C.c_int), ("action_f64", C.c_int),
                ("nx", C.c_int), ("nt", C.c_int), ("substeps", C.c_int), ("control_neumann", C.c_int),
                ("normalize", C.c_int), ("dt", C.c_double), ("dx", C.c_double), ("max_control_value", C.c_double)]


class ArzCfg(C.Structure):
    _fields_ = [("M", C.c_int), ("control_freq", C.c_int), ("sim_type", C.c_int), ("action_dim", C.c_int),
                ("T", C.c_double), ("dt", C.c_double), ("dx", C.c_double), ("tau", C.c_double), ("vm", C.c_double), ("rm", C.c_double)]


class NsCfg(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("n_iter", C.c_int), ("bc", (C.c_int * 2) * 4),
                ("dt", C.c_double), ("dx", C.c_double), ("dy", C.c_double), ("nu", C.c_double), ("rho", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_SO)
        _lib.pdecg_oracle_bc1d_step.argtypes = [C.POINTER(Cfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        _lib.pdecg_oracle_bc1d_step.restype = C.c_int
        _lib.pdecg_oracle_arz_step.argtypes = [C.POINTER(ArzCfg), C.c_int64] + [C.c_void_p] * 6
        _lib.pdecg_oracle_arz_step.restype = C.c_int
        _lib.pdecg_oracle_ns_step.argtypes = [C.POINTER(NsCfg), C.c_int64] + [C.c_void_p] * 4
        _lib.pdecg_oracle_ns_step.restype = C.c_int
    return _lib


class Bc1dCOracle:
    """Batched state integrator: `step(action)` advances `u` in place by one control step."""

    def __init__(self, env, T, dt, X, dx, control_sample_rate, control_type="Dirchilet", normalize=False,
                 max_control_value=20, precision="ref"):
        self.cfg = Cfg()
        self.cfg.env = 1 if env == "rd" else 0
        self.cfg.state_f64 = int(precision == "fp64")
        self.cfg.nx = int(round(X / dx))
        self.cfg.nt = int(round(T / dt) + 1)
        self.cfg.substeps = int(round(control_sample_rate / dt))
        self.cfg.control_neumann = int(control_type == "Neumann")
        self.cfg.normalize = int(bool(normalize))
        self.cfg.dt, self.cfg.dx, self.cfg.max_control_value = float(dt), float(dx), float(max_control_value)
        self.npts = self.cfg.nx + (1 if env == "rd" else 0)

    def reset(self, init_condition, beta):
        sd = np.float64 if self.cfg.state_f64 else np.float32
        ic = np.asarray(init_condition)
        self.u = np.ascontiguousarray(ic.astype(sd) if ic.ndim == 2 else np.broadcast_to(ic.astype(sd), (1, self.npts)))
        self.n = self.u.shape[0]
        b = np.asarray(beta)
        if self.cfg.state_f64:
            b = b.astype(np.float64)
        self.beta = np.ascontiguousarray(b)
        self.cfg.beta_f64 = int(self.beta.dtype == np.float64)
        self.t = np.zeros(self.n, dtype=np.int32)
        return self.u

    def step(self, action):
        a = np.ascontiguousarray(action).reshape(self.n)
        assert a.dtype in (np.float32, np.float64)
        self.cfg.action_f64 = int(a.dtype == np.float64)
        stride = 0 if self.beta.ndim == 1 else self.npts
        lib().pdecg_oracle_bc1d_step(C.byref(self.cfg), self.n, self.u.ctypes.data, self.beta.ctypes.data, stride,
                                     a.ctypes.data, self.t.ctypes.data)
        return self.u


class ArzCOracle:
    """TrafficPDE1D state integrator (traffic_arz_env.py:131-233) in C: `step(action)` advances r, y, v, time in place.
    Initial rows / steady state come from the caller (e.g. a numpy-oracle instance after reset())."""
    _TYPES = {"inlet": 0, "outlet": 1, "both": 2, "outlet-train": 3}

    def __init__(self, T, dt, X, dx, simulation_type="inlet", v_max=40, ro_max=0.16, tau=60, control_freq=1):
        self.cfg = ArzCfg()
        self.cfg.M = len(np.arange(0, X + dx, dx))
        self.cfg.control_freq, self.cfg.sim_type = control_freq, self._TYPES[simulation_type]
        self.cfg.action_dim = 2 if simulation_type == "both" else 1
        self.cfg.T, self.cfg.dt, self.cfg.dx, self.cfg.tau = float(T), float(dt), float(dx), float(tau)
        self.cfg.vm, self.cfg.rm = float(v_max), float(ro_max)

    def reset(self, r, y, qs, time=None):
        self.r, self.y = np.ascontiguousarray(r, dtype=np.float64).copy(), np.ascontiguousarray(y, dtype=np.float64).copy()
        self.n = self.r.shape[0]
        self.v = np.zeros_like(self.r)
        self.qs = np.ascontiguousarray(np.broadcast_to(np.asarray(qs, dtype=np.float64), (self.n,))).copy()
        self.time = np.zeros(self.n) if time is None else np.ascontiguousarray(time, dtype=np.float64).copy()

    def step(self, action):
        a = np.ascontiguousarray(np.asarray(action, dtype=np.float64).reshape(self.n, self.cfg.action_dim))
        lib().pdecg_oracle_arz_step(C.byref(self.cfg), self.n, self.r.ctypes.data, self.y.ctypes.data, self.v.ctypes.data,
                                    a.ctypes.data, self.qs.ctypes.data, self.time.ctypes.data)
        return np.concatenate((self.r, self.v), axis=1)


class NsCOracle:
    """NavierStokes2D step (navier_stokes2D.py:118-157) in C: `step(action)` advances u, v, p in place."""
    _BC = {"Dirchilet": 0, "Neumann": 1, "Controllable": 2}

    def __init__(self, dt, dx, dy, nx, ny, boundary_condition, viscosity=0.1, density=1.0, maximum_pressure_iteration=2000):
        self.cfg = NsCfg()
        self.cfg.nx, self.cfg.ny, self.cfg.n_iter = nx, ny, int(maximum_pressure_iteration)
        for s, side in enumerate(("lower", "upper", "left", "right")):
            for c in range(2):
                self.cfg.bc[s][c] = self._BC[boundary_condition[side][c]]
        self.cfg.dt, self.cfg.dx, self.cfg.dy = float(dt), float(dx), float(dy)
        self.cfg.nu, self.cfg.rho = float(viscosity), float(density)

    def reset(self, u, v, p):
        self.u, self.v, self.p = (np.ascontiguousarray(f, dtype=np.float64).copy() for f in (u, v, p))
        self.n = self.u.shape[0]

    def step(self, action):
        a = np.ascontiguousarray(np.asarray(action, dtype=np.float64).reshape(self.n))
        lib().pdecg_oracle_ns_step(C.byref(self.cfg), self.n, self.u.ctypes.data, self.v.ctypes.data, self.p.ctypes.data, a.ctypes.data)
        return np.stack((self.u, self.v), axis=-1)

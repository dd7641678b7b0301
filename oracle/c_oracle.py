"""ctypes wrapper of oracle/c/pdecg_oracle.c (plain-C restatement of the Transport / ReactionDiffusion
integrators) -- TEST INFRASTRUCTURE ONLY; see the C file's header.  State-only (no rewards): it checks the
integrators at full batch sizes and times a compiled multi-threaded CPU port."""
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


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_SO)
        _lib.pdecg_oracle_bc1d_step.argtypes = [C.POINTER(Cfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        _lib.pdecg_oracle_bc1d_step.restype = C.c_int
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
        assert self.npts <= 520

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

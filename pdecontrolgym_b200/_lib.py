"""ctypes binding of libpdecg.so (the C ABI declared in include/pdecg.h).

There is deliberately no fallback: if the shared library is missing or was not built for this
machine, importing an env raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C pdecontrolgym_b200/csrc`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PDECG_LIB", os.path.join(_HERE, "lib", "libpdecg.so"))   # override: A/B testing of builds

ABI_VERSION = 2
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA = 0, -1, -2, -3
PREC = {"ref": 0, "fp64": 1, "fp32": 2}
F32, F64 = 0, 1


class PdecgError(RuntimeError):
    """A libpdecg entry point returned a negative status."""

    def __init__(self, code, msg):
        super().__init__(f"libpdecg error {code}: {msg}")
        self.code = code


class Bc1dCfg(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("env", C.c_int32), ("precision", C.c_int32),
        ("beta_dtype", C.c_int32), ("action_dtype", C.c_int32), ("nx", C.c_int32),
        ("nt", C.c_int32), ("substeps", C.c_int32), ("dt", C.c_double), ("dx", C.c_double),
        ("control_type", C.c_int32), ("sensing_loc", C.c_int32), ("sensing_type", C.c_int32),
        ("normalize", C.c_int32), ("max_control_value", C.c_double),
        ("limit_pde_state_size", C.c_int32), ("reward_kind", C.c_int32),
        ("max_state_value", C.c_double), ("reward_nt", C.c_int32), ("norm_ord", C.c_int32),
        ("horizon", C.c_int32), ("t_horizon_length", C.c_int32),
        ("truncate_penalty", C.c_double), ("terminate_reward", C.c_double), ("norm_coeff", C.c_double),
        ("truncate_nonfinite", C.c_int32),
    ]


class Bc1dBuffers(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int64), ("u", C.c_void_p), ("beta", C.c_void_p), ("beta_env_stride", C.c_int64),
        ("action", C.c_void_p), ("obs", C.c_void_p), ("reward", C.c_void_p),
        ("terminated", C.c_void_p), ("truncated", C.c_void_p), ("time_index", C.c_void_p),
        ("carry", C.c_void_p), ("obs_copy", C.c_void_p), ("obs_copy_f32", C.c_int32), ("nonfinite", C.c_void_p),
    ]


class Bc1dAutoReset(C.Structure):
    _fields_ = [
        ("init", C.c_void_p), ("init_dtype", C.c_int32), ("random", C.c_int32), ("init_env_stride", C.c_int64),
        ("lo", C.c_double), ("hi", C.c_double), ("seed", C.c_uint64), ("env_id_offset", C.c_int64),
        ("episode", C.c_void_p), ("final_obs", C.c_void_p),
    ]


class BtCfg(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("nx", C.c_int32), ("nt", C.c_int32), ("weekends", C.c_int32),
        ("dt", C.c_double), ("dx", C.c_double), ("k", C.c_double), ("t1_threshold_k", C.c_double),
        ("t2_threshold_k", C.c_double), ("dosage_termination_threshold", C.c_double),
        ("alpha", C.c_double), ("alpha_beta_ratio", C.c_double), ("t1_detection_radius", C.c_double),
        ("t1_death_radius", C.c_double), ("total_dosage", C.c_double), ("precision", C.c_int32),
    ]


class BtBuffers(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int64), ("u", C.c_void_p), ("obs", C.c_void_p), ("x_scale", C.c_void_p),
        ("action", C.c_void_p), ("ienv", C.c_void_p), ("fenv", C.c_void_p), ("reward", C.c_void_p),
        ("terminated", C.c_void_p), ("truncated", C.c_void_p), ("valid", C.c_void_p),
        ("final_obs", C.c_void_p), ("mask", C.c_void_p), ("stage_out", C.c_void_p),
    ]


class ArzCfg(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("simulation_type", C.c_int32), ("M", C.c_int32), ("control_freq", C.c_int32),
        ("limit_pde_state_size", C.c_int32), ("action_dim", C.c_int32), ("T", C.c_double), ("dt", C.c_double),
        ("t_over_dt", C.c_double), ("vm", C.c_double), ("rm", C.c_double), ("c1", C.c_double), ("c2", C.c_double),
        ("c3", C.c_double), ("c4", C.c_double), ("vm_over_rm", C.c_double), ("precision", C.c_int32),
    ]


class ArzBuffers(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int64), ("ry", C.c_void_p), ("obs", C.c_void_p), ("action", C.c_void_p),
        ("fenv", C.c_void_p), ("reward", C.c_void_p), ("terminated", C.c_void_p), ("truncated", C.c_void_p),
        ("final_obs", C.c_void_p),
    ]


class NsCfg(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("nt", C.c_int32),
        ("n_pressure_iters", C.c_int32), ("bc", (C.c_int32 * 2) * 4), ("n_action_ref", C.c_int32),
        ("n_uref", C.c_int32), ("dt", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
        ("viscosity", C.c_double), ("density", C.c_double), ("gamma", C.c_double),
        ("precision", C.c_int32), ("reserved", C.c_int32),
    ]


class NsBuffers(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int64), ("u", C.c_void_p), ("v", C.c_void_p), ("p", C.c_void_p), ("obs", C.c_void_p),
        ("action", C.c_void_p), ("u_ref", C.c_void_p), ("action_ref", C.c_void_p), ("time_index", C.c_void_p),
        ("reward", C.c_void_p), ("terminated", C.c_void_p), ("truncated", C.c_void_p), ("valid", C.c_void_p),
        ("final_obs", C.c_void_p), ("work", C.c_void_p),
    ]


class NeuronCfg(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("M", C.c_int32), ("nt", C.c_int32), ("reserved", C.c_int32), ("dt", C.c_double),
        ("dx", C.c_double), ("D", C.c_double), ("a", C.c_double), ("g", C.c_double), ("RSubG", C.c_double),
        ("aTilde1", C.c_double), ("beta", C.c_double), ("LSubS", C.c_double), ("inv_scale", C.c_double),
        ("H0", C.c_double), ("H1", C.c_double), ("coef", C.c_double),
    ]


class NeuronBuffers(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int64), ("u", C.c_void_p), ("Z", C.c_void_p), ("L", C.c_void_p), ("time", C.c_void_p),
        ("control", C.c_void_p), ("pB", C.c_void_p), ("p", C.c_void_p), ("terminated", C.c_void_p),
    ]


_lib = None


def lib():
    """Load (once) and return the ctypes handle of libpdecg.so; raise loudly if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension has not been built. "
            "Run `make -C pdecontrolgym_b200/csrc` (needs nvcc; no GPU required). "
            "pdecontrolgym_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.pdecg_last_error.restype = C.c_char_p
    L.pdecg_abi_version.restype = C.c_int32
    L.pdecg_launch_count.restype = C.c_int64
    L.pdecg_bc1d_npts.argtypes = [C.POINTER(Bc1dCfg)]
    L.pdecg_bc1d_npts.restype = C.c_int32
    L.pdecg_bc1d_carry_len.argtypes = [C.POINTER(Bc1dCfg)]
    L.pdecg_bc1d_carry_len.restype = C.c_int32
    L.pdecg_bc1d_step.argtypes = [C.POINTER(Bc1dCfg), C.POINTER(Bc1dBuffers), C.c_void_p]
    L.pdecg_bc1d_step.restype = C.c_int
    L.pdecg_bc1d_step_autoreset.argtypes = [C.POINTER(Bc1dCfg), C.POINTER(Bc1dBuffers), C.POINTER(Bc1dAutoReset), C.c_void_p]
    L.pdecg_bc1d_step_autoreset.restype = C.c_int
    L.pdecg_bc1d_reset.argtypes = [C.POINTER(Bc1dCfg), C.POINTER(Bc1dBuffers), C.c_void_p, C.c_int32,
                                   C.c_int64, C.c_void_p, C.c_void_p]
    L.pdecg_bc1d_reset.restype = C.c_int
    L.pdecg_bc1d_reset_random.argtypes = [C.POINTER(Bc1dCfg), C.POINTER(Bc1dBuffers), C.c_void_p, C.c_int32, C.c_double,
                                          C.c_double, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pdecg_bc1d_reset_random.restype = C.c_int
    L.pdecg_bc1d_step_host.argtypes = [C.POINTER(Bc1dCfg), C.POINTER(Bc1dBuffers), C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Bc1dAutoReset), C.c_void_p]
    L.pdecg_bc1d_step_host.restype = C.c_int
    L.pdecg_bt_reset.argtypes = [C.POINTER(BtCfg), C.POINTER(BtBuffers), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.pdecg_bt_reset.restype = C.c_int
    L.pdecg_bt_run.argtypes = [C.POINTER(BtCfg), C.POINTER(BtBuffers), C.c_int32, C.c_void_p]
    L.pdecg_bt_run.restype = C.c_int
    L.pdecg_bt_step_autoreset.argtypes = [C.POINTER(BtCfg), C.POINTER(BtBuffers), C.c_void_p, C.c_int64, C.c_void_p]
    L.pdecg_bt_step_autoreset.restype = C.c_int
    L.pdecg_bt_step_host.argtypes = [C.POINTER(BtCfg), C.POINTER(BtBuffers)] + [C.c_void_p] * 7 + [C.c_int64, C.c_void_p]
    L.pdecg_bt_step_host.restype = C.c_int
    L.pdecg_arz_reset.argtypes = [C.POINTER(ArzCfg), C.POINTER(ArzBuffers), C.c_void_p, C.c_void_p, C.c_void_p]
    L.pdecg_arz_reset.restype = C.c_int
    L.pdecg_arz_step.argtypes = [C.POINTER(ArzCfg), C.POINTER(ArzBuffers), C.c_void_p]
    L.pdecg_arz_step.restype = C.c_int
    L.pdecg_arz_step_autoreset.argtypes = [C.POINTER(ArzCfg), C.POINTER(ArzBuffers), C.c_void_p, C.c_void_p]
    L.pdecg_arz_step_autoreset.restype = C.c_int
    L.pdecg_arz_step_host.argtypes = [C.POINTER(ArzCfg), C.POINTER(ArzBuffers)] + [C.c_void_p] * 7
    L.pdecg_arz_step_host.restype = C.c_int
    L.pdecg_ns_reset.argtypes = [C.POINTER(NsCfg), C.POINTER(NsBuffers), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_void_p]
    L.pdecg_ns_reset.restype = C.c_int
    L.pdecg_ns_step.argtypes = [C.POINTER(NsCfg), C.POINTER(NsBuffers), C.c_void_p]
    L.pdecg_ns_step.restype = C.c_int
    L.pdecg_ns_step_autoreset.argtypes = [C.POINTER(NsCfg), C.POINTER(NsBuffers), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_void_p]
    L.pdecg_ns_step_autoreset.restype = C.c_int
    L.pdecg_ns_step_host.argtypes = [C.POINTER(NsCfg), C.POINTER(NsBuffers)] + [C.c_void_p] * 9 + [C.c_int64, C.c_void_p]
    L.pdecg_ns_step_host.restype = C.c_int
    L.pdecg_ns_solve_pressure.argtypes = [C.POINTER(NsCfg), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pdecg_ns_solve_pressure.restype = C.c_int
    L.pdecg_ns_workspace_bytes.argtypes = [C.POINTER(NsCfg), C.c_int64]
    L.pdecg_ns_workspace_bytes.restype = C.c_size_t
    L.pdecg_selftest_constdiv.argtypes = [C.c_int32, C.c_int64, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pdecg_selftest_constdiv.restype = C.c_int
    L.pdecg_neuron_step.argtypes = [C.POINTER(NeuronCfg), C.POINTER(NeuronBuffers), C.c_int32, C.c_void_p]
    L.pdecg_neuron_step.restype = C.c_int
    L.pdecg_selftest_fastdiv.argtypes = [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pdecg_selftest_fastdiv.restype = C.c_int
    L.pdecg_selftest_exp.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pdecg_selftest_exp.restype = C.c_int
    L.pdecg_probe_fp_issue.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float),
                                       C.POINTER(C.c_double), C.c_void_p]
    L.pdecg_probe_fp_issue.restype = C.c_int
    if L.pdecg_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libpdecg ABI {L.pdecg_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise PdecgError(rc, lib().pdecg_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib().pdecg_launch_count())


class _NoCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NOCTX = _NoCtx()


def device_guard(device):
    """`with device_guard(dev):` == `with torch.cuda.device(dev):`, minus the switch (and its ~5 us of host time per
    step) when `dev` is already the current device -- the normal case of one process per GPU."""
    import torch
    return _NOCTX if torch.cuda.current_device() == device.index else torch.cuda.device(device)


def current_raw_stream(device_index: int) -> int:
    """cudaStream_t of torch's current stream on `device_index` (the fast private accessor when available)."""
    import torch
    try:
        return torch._C._cuda_getCurrentRawStream(device_index)
    except AttributeError:  # pragma: no cover - older/newer torch
        return torch.cuda.current_stream(device_index).cuda_stream


def host_ptr(t, dtype, numel, what):
    """data_ptr() of an optional contiguous CPU tensor argument of a `*_step_host` call (None -> NULL)."""
    if t is None:
        return None
    if t.device.type != "cpu" or not t.is_contiguous() or t.dtype != dtype or t.numel() != numel:
        raise ValueError(f"{what} must be a contiguous CPU {dtype} tensor with {numel} elements")
    return t.data_ptr()

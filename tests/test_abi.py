"""CPU: the C-ABI shared library loads, exports every function include/pdecg.h declares, and the ctypes
mirrors of the structs have the sizes the C compiler gives them.  No compute call is made (no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pdecg.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pdecg_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from pdecontrolgym_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib


def test_header_declares_the_expected_entry_points():
    fns = declared_functions()
    for must in ("pdecg_bc1d_step", "pdecg_bc1d_reset", "pdecg_bc1d_step_host", "pdecg_bt_run", "pdecg_bt_reset",
                 "pdecg_arz_step", "pdecg_arz_reset", "pdecg_ns_step", "pdecg_ns_reset", "pdecg_ns_solve_pressure",
                 "pdecg_last_error", "pdecg_abi_version", "pdecg_launch_count"):
        assert must in fns


def test_library_exports_every_declared_symbol(lib):
    L = lib.lib()
    missing = [f for f in declared_functions() if not hasattr(L, f)]
    assert not missing, f"libpdecg.so lacks {missing}"
    assert L.pdecg_abi_version() == lib.ABI_VERSION
    assert L.pdecg_launch_count() == 0          # nothing was launched by loading / symbol lookup


def test_struct_layouts_match_the_c_compiler(lib, tmp_path):
    names = [("pdecg_bc1d_cfg", lib.Bc1dCfg), ("pdecg_bc1d_buffers", lib.Bc1dBuffers), ("pdecg_bt_cfg", lib.BtCfg),
             ("pdecg_bt_buffers", lib.BtBuffers), ("pdecg_arz_cfg", lib.ArzCfg), ("pdecg_arz_buffers", lib.ArzBuffers),
             ("pdecg_ns_cfg", lib.NsCfg), ("pdecg_ns_buffers", lib.NsBuffers)]
    src = tmp_path / "sz.c"
    body = "\n".join(f'  printf("%zu\\n", sizeof({c}));' for c, _ in names)
    # a few field offsets as well
    body += '\n  printf("%zu\\n", offsetof(pdecg_bc1d_cfg, truncate_penalty));'
    body += '\n  printf("%zu\\n", offsetof(pdecg_ns_cfg, dt));'
    body += '\n  printf("%zu\\n", offsetof(pdecg_arz_cfg, t_over_dt));'
    src.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "pdecg.h"\nint main(void){{\n{body}\n  return 0;}}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(x) for x in out]
    for (cname, cls), sz in zip(names, sizes):
        assert ctypes.sizeof(cls) == sz, f"{cname}: ctypes {ctypes.sizeof(cls)} vs C {sz}"
    assert lib.Bc1dCfg.truncate_penalty.offset == sizes[len(names)]
    assert lib.NsCfg.dt.offset == sizes[len(names) + 1]
    assert lib.ArzCfg.t_over_dt.offset == sizes[len(names) + 2]


def test_invalid_configs_are_rejected_without_a_gpu(lib):
    """Argument validation happens before any CUDA call, so it is testable here."""
    L = lib.lib()
    cfg = lib.Bc1dCfg()
    buf = lib.Bc1dBuffers()
    assert L.pdecg_bc1d_step(ctypes.byref(cfg), ctypes.byref(buf), None) == lib.ERR_INVALID      # abi_version 0
    assert b"abi_version" in L.pdecg_last_error()
    cfg.abi_version = lib.ABI_VERSION
    cfg.env, cfg.nx, cfg.nt, cfg.substeps, cfg.dt, cfg.dx = 0, 100, 10001, 30, 1e-4, 1e-2
    cfg.reward_kind, cfg.reward_nt = 0, 10000
    assert lib.lib().pdecg_bc1d_carry_len(ctypes.byref(cfg)) == 102      # 30 does not divide 100: general 101-entry ring
    cfg.nt = 50
    assert L.pdecg_bc1d_step(ctypes.byref(cfg), ctypes.byref(buf), None) == lib.ERR_UNSUPPORTED  # look-back wraps onto written rows
    assert b"TunedReward1D" in L.pdecg_last_error()
    cfg.nt = 10001
    cfg.substeps = 1000
    cfg.nx = 600                 # beyond the warp kernels' 512 points: served by the CTA-per-env kernel
    assert L.pdecg_bc1d_step(ctypes.byref(cfg), ctypes.byref(buf), None) == lib.OK               # n_envs == 0: nothing to launch
    cfg.nx = 8193
    assert L.pdecg_bc1d_step(ctypes.byref(cfg), ctypes.byref(buf), None) == lib.ERR_UNSUPPORTED
    assert b"shared-memory-resident" in L.pdecg_last_error()
    acfg = lib.ArzCfg()
    acfg.abi_version, acfg.simulation_type, acfg.M, acfg.control_freq, acfg.action_dim = lib.ABI_VERSION, 7, 51, 1, 1
    assert L.pdecg_arz_step(ctypes.byref(acfg), ctypes.byref(lib.ArzBuffers()), None) == lib.ERR_INVALID
    assert b"Invalid simulation type" in L.pdecg_last_error()
    # fused auto-reset entry points: the restart source is mandatory; grid limits are reported before any launch
    acfg.simulation_type, acfg.M = 1, 51
    abuf = lib.ArzBuffers()
    assert L.pdecg_arz_step_autoreset(ctypes.byref(acfg), ctypes.byref(abuf), None, None) == lib.ERR_INVALID
    assert b"profile" in L.pdecg_last_error()
    acfg.M = 513
    assert L.pdecg_arz_step(ctypes.byref(acfg), ctypes.byref(abuf), None) == lib.ERR_UNSUPPORTED
    bcfg = lib.BtCfg()
    bcfg.abi_version, bcfg.nx, bcfg.nt, bcfg.dt, bcfg.dx, bcfg.k = lib.ABI_VERSION, 201, 601, 1.0, 1.0, 1e5
    assert L.pdecg_bt_step_autoreset(ctypes.byref(bcfg), ctypes.byref(lib.BtBuffers()), None, 0, None) == lib.ERR_INVALID
    assert b"init" in L.pdecg_last_error()
    bcfg.nx = 513
    assert L.pdecg_bt_run(ctypes.byref(bcfg), ctypes.byref(lib.BtBuffers()), 0, None) == lib.ERR_UNSUPPORTED
    ncfg = lib.NsCfg()
    ncfg.abi_version, ncfg.nx, ncfg.ny, ncfg.nt, ncfg.n_pressure_iters = lib.ABI_VERSION, 21, 21, 200, 10
    ncfg.dt, ncfg.dx, ncfg.dy = 1e-3, 0.05, 0.05
    assert L.pdecg_ns_step_autoreset(ctypes.byref(ncfg), ctypes.byref(lib.NsBuffers()), None, None, None, 0, None) == lib.ERR_INVALID
    assert b"init" in L.pdecg_last_error()
    assert L.pdecg_ns_step(ctypes.byref(ncfg), ctypes.byref(lib.NsBuffers()), None) == lib.OK          # n_envs == 0
    ncfg.nx = ncfg.ny = 64
    assert L.pdecg_ns_step(ctypes.byref(ncfg), ctypes.byref(lib.NsBuffers()), None) == lib.OK          # 4,096 points still fit
    assert L.pdecg_ns_workspace_bytes(ctypes.byref(ncfg), 10) == 0                                      # lives on chip
    ncfg.nx = ncfg.ny = 65                                                                             # HBM-resident path: needs scratch
    assert L.pdecg_ns_workspace_bytes(ctypes.byref(ncfg), 10) == 10 * (2 * 65 * 65 + 3 * 76 * 76) * 8 + 256
    assert L.pdecg_ns_step(ctypes.byref(ncfg), ctypes.byref(lib.NsBuffers()), None) == lib.OK          # n_envs == 0
    ncfg.nx = ncfg.ny = 8193
    assert L.pdecg_ns_step(ctypes.byref(ncfg), ctypes.byref(lib.NsBuffers()), None) == lib.ERR_UNSUPPORTED
    assert b"exceeds the limit" in L.pdecg_last_error()


def test_no_cpu_fallback():
    """Without a CUDA device the envs refuse to construct -- the oracle is never used as a fallback."""
    import torch
    import pdecontrolgym_b200 as pcg
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(RuntimeError):
        pcg.make("PDEControlGym-TransportPDE1D", num_envs=2, device="cpu", T=1, dt=1e-4, X=1, dx=1e-2,
                 reward_class=pcg.TunedReward1D(10000))
    with pytest.raises(Exception):
        pcg.make("PDEControlGym-TransportPDE1D", num_envs=2, device="cuda", T=1, dt=1e-4, X=1, dx=1e-2,
                 reward_class=pcg.TunedReward1D(10000))
    src = "".join(open(os.path.join(ROOT, "pdecontrolgym_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "pdecontrolgym_b200")) if f.endswith(".py"))
    assert "oracle" not in src.replace("oracle/", "").lower() or "import oracle" not in src
    assert "from oracle" not in src and "import oracle" not in src

"""Does an FP64 (FP32) warp instruction with 16 active lanes occupy its pipe for less time than one with 32?  (No: the
warp-instruction rate per SM is the same, so a half-empty warp wastes half of its pipe time.)  python scripts/probe_half_warp.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pdecontrolgym_b200 import _lib
L = _lib.lib()
L.pdecg_probe_fp_issue.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_void_p]
torch.zeros(1, device="cuda:0")
for dtype, name in ((_lib.F64, "f64"), (_lib.F32, "f32")):
    for threads, blocks in ((32, 148 * 16), (16, 148 * 16), (16, 148 * 32), (32, 148 * 4), (16, 148 * 4), (16, 148 * 8)):
        ms, ops = C.c_float(), C.c_double()
        rc = L.pdecg_probe_fp_issue(dtype, blocks, threads, 200000, C.byref(ms), C.byref(ops), None)
        assert rc == 0, L.pdecg_last_error()
        print(name, "threads/block", threads, "blocks", blocks, "blocks/SM", blocks // 148, f"{ms.value:.3f} ms  {ops.value / ms.value / 1e9:.2f} Tlane-op/s  warp-instr/s/SM {ops.value/threads/ms.value/1e3/148:.3e}")

"""What the box can move device -> pinned host when N GPUs copy at once: the ceiling of the end-to-end (`e2e`) leg of
bench.py, whose step ships every env's observation row to the host.  Run under torchrun with N = 1, 2, 4, 8 ranks; rank 0
prints one JSON line per variant with the aggregate GB/s (bytes of all ranks / max-over-ranks time).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/d2h_ceiling.py [--mb 53]

Variants: how the host buffer is allocated (torch pin_memory, cudaHostAlloc default / write-combined, 2 MiB-aligned
mmap + MADV_HUGEPAGE + cudaHostRegister) x how the copy is issued (one stream, or 4 streams x quarter-size chunks).
The GPU boxes of this pool are KVM guests with a single virtual NUMA node (no numactl, no per-GPU affinity to bind
to), so NUMA placement is not a knob here -- the script prints what `nvidia-smi topo -m` and /sys say about it.
"""
import argparse
import ctypes as C
import json
import mmap
import os
import subprocess
import sys

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=53.3, help="bytes per rank per copy (MB): one step's float32 observations")
    ap.add_argument("--reps", type=int, default=30)
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = int(args.mb * 1e6) // 4096 * 4096
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev).random_(0, 255)
    rt = C.CDLL("libcudart.so.12")
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaFreeHost.argtypes = [C.c_void_p]
    rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
    rt.cudaHostUnregister.argtypes = [C.c_void_p]
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    streams = [torch.cuda.Stream(dev) for _ in range(4)]
    D2H, H2D = 2, 1

    def run(name, host_ptr, nstreams, kind=D2H):
        chunk = nbytes // nstreams // 4096 * 4096

        def once():
            for s in range(nstreams):
                off = s * chunk
                a, b = (host_ptr + off, src.data_ptr() + off) if kind == D2H else (src.data_ptr() + off, host_ptr + off)
                rc = rt.cudaMemcpyAsync(a, b, chunk, kind, C.c_void_p(streams[s].cuda_stream))
                assert rc == 0, rc
        for _ in range(3):
            once()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(nstreams)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(nstreams)]
        for s in range(nstreams):
            ev0[s].record(streams[s])
        for _ in range(args.reps):
            once()
        for s in range(nstreams):
            ev1[s].record(streams[s])
        torch.cuda.synchronize(dev)
        ms = max(ev0[0].elapsed_time(e) for e in ev1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            gbs = world * chunk * nstreams * args.reps / (float(t.item()) * 1e-3) / 1e9
            print(json.dumps({"n_gpus": world, "variant": name, "direction": "d2h" if kind == D2H else "h2d", "streams": nstreams,
                              "mb_per_rank": chunk * nstreams / 1e6, "aggregate_gbs": round(gbs, 2), "per_gpu_gbs": round(gbs / world, 2)}), flush=True)

    # 1. torch's pinned allocator
    t_pin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    run("torch pin_memory", t_pin.data_ptr(), 1)
    run("torch pin_memory", t_pin.data_ptr(), 4)
    run("torch pin_memory", t_pin.data_ptr(), 1, H2D)
    # 2. cudaHostAlloc, default and write-combined
    for flags, name in ((0, "cudaHostAlloc default"), (4, "cudaHostAlloc write-combined")):
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), nbytes, flags) == 0
        run(name, p.value, 1)
        run(name, p.value, 4)
        rt.cudaFreeHost(p)
    # 3. transparent huge pages + cudaHostRegister
    try:
        mm = mmap.mmap(-1, nbytes + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        base = C.addressof(C.c_char.from_buffer(mm))
        aligned = (base + (2 << 20) - 1) // (2 << 20) * (2 << 20)
        libc = C.CDLL("libc.so.6", use_errno=True)
        libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        libc.madvise(C.c_void_p(aligned), nbytes, 14)          # MADV_HUGEPAGE
        C.memset(C.c_void_p(aligned), 1, nbytes)                # touch
        assert rt.cudaHostRegister(C.c_void_p(aligned), nbytes, 0) == 0
        run("mmap + MADV_HUGEPAGE + cudaHostRegister", aligned, 1)
        run("mmap + MADV_HUGEPAGE + cudaHostRegister", aligned, 4)
        rt.cudaHostUnregister(C.c_void_p(aligned))
    except Exception as exc:  # noqa: BLE001
        if rank == 0:
            print(json.dumps({"n_gpus": world, "variant": "hugepage", "error": repr(exc)[:200]}), flush=True)
    if rank == 0:
        def sh(cmd):
            try:
                return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
            except Exception as exc:  # noqa: BLE001
                return repr(exc)
        print(json.dumps({"n_gpus": world, "host": {"cpus": len(os.sched_getaffinity(0)), "numa_nodes": sh("ls -d /sys/devices/system/node/node* | wc -l"),
                                                     "hypervisor": sh("lscpu | grep -i 'Hypervisor vendor'"), "thp": sh("cat /sys/kernel/mm/transparent_hugepage/enabled"),
                                                     "mem_gb": sh("awk '/MemTotal/{print int($2/1e6)}' /proc/meminfo")}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())

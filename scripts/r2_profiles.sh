#!/bin/bash
# Round-2 evidence in one gpurun call (one GPU): the launch list of a whole bench.py run, `ncu --set full` captures of the
# kernels this round touched, the all-env table and the large-grid NavierStokes bench.  Outputs land in gpurun_out/r2_*.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
bash scripts/ncu_capture.sh r2 ns:ns_step_warp:2 ns_tiled:ns_lg_jacobi:3 arz:arz_step_kernel:2 bt:bt_run_kernel:2
python scripts/bench_ns_large.py > gpurun_out/r2_ns_large.jsonl 2>&1
python scripts/bench_all_envs.py > gpurun_out/r2_all_envs.jsonl 2>&1
tail -3 gpurun_out/r2_all_envs.jsonl

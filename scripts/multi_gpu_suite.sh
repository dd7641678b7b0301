#!/bin/bash
# Everything that needs more than one GPU, in one gpurun call (`gpurun --gpus 8 -- bash scripts/multi_gpu_suite.sh`):
#   1. the box's concurrent device->host ceiling at N = 1, 2, 4, 8 (scripts/d2h_ceiling.py)
#   2. bench.py at N = 2, 4, 8 (headline weak scaling + e2e with its in-run D2H ceiling + config 5 strong scaling)
#   3. config 5 weak scaling: 2^20 envs per GPU at N = 1 and 8 (scripts/ppo_rollout.py)
#   4. the two-rank sharding parity test on two distinct GPUs
mkdir -p gpurun_out
out=gpurun_out/r2_multi
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > ${out}_topo.txt 2>&1
: > ${out}_d2h_ceiling.jsonl
for n in 1 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29500 + n)) scripts/d2h_ceiling.py 2>> ${out}_err.log | grep '^{' >> ${out}_d2h_ceiling.jsonl
done
for n in 2 4 8; do
  timeout 400 $TR --nproc-per-node $n --master-port $((29520 + n)) bench.py --gpus $n --steps 20 --warmup 5 2>> ${out}_err.log | grep '^{' > ${out}_bench_n$n.json
done
: > ${out}_ppo_weak.jsonl
for n in 1 8; do
  for e in rd transport; do
    timeout 300 $TR --nproc-per-node $n --master-port $((29540 + n)) scripts/ppo_rollout.py --env $e --envs $((n << 20)) --rollout 16 --repeats 2 2>> ${out}_err.log | grep '^{' >> ${out}_ppo_weak.jsonl
  done
done
timeout 300 python -m pytest tests/test_gpu_sharding.py -m gpu -q 2>&1 | tail -3 > ${out}_sharding_test.log
cat ${out}_sharding_test.log
wc -l ${out}_d2h_ceiling.jsonl ${out}_ppo_weak.jsonl
tail -c 600 ${out}_err.log

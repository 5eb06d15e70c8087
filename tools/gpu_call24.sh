#!/bin/bash
# eight GPUs of one node: the bench line at N = 8 (replicas, pixel-sharded leg, parity, config #5)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c24_gpus.log 2>&1
echo "== bench N=8"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c24_bench_n8.json 2> gpurun_out/c24_bench_n8.err; tail -c 2500 gpurun_out/c24_bench_n8.json; tail -5 gpurun_out/c24_bench_n8.err

#!/bin/bash
# two GPUs: the pixel-sharded mode against one GPU, then the bench line at N = 2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c23_gpus.log 2>&1
PYT="python -m pytest -q -m gpu --timeout 150 --timeout-method=thread -p no:cacheprovider"
echo "== multi"; timeout 500 $PYT tests/test_gpu_multi.py > gpurun_out/c23_multi.log 2>&1; tail -15 gpurun_out/c23_multi.log
echo "== bench N=2"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c23_bench_n2.json 2> gpurun_out/c23_bench_n2.err; tail -c 2500 gpurun_out/c23_bench_n2.json; tail -5 gpurun_out/c23_bench_n2.err

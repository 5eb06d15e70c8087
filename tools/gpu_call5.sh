#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 90 --timeout-method=thread -p no:cacheprovider"
echo "== T1+T2 quick"; timeout 400 $PYT tests/test_gpu_kernels.py tests/test_gpu_render.py -k "(hashgrid and not cell) or t2_same_stream or merge_group or path_length or full_size_same" > gpurun_out/c5_t.log 2>&1; tail -4 gpurun_out/c5_t.log
echo "== variants"
timeout 400 python tools/gpu_variants.py \
  lane=:SVCM_MERGE_KERNEL=1 \
  lane_serial=:SVCM_MERGE_KERNEL=1,SVCM_NO_PASS_OVERLAP=1 \
  c2=:W=512,ITERS=40 \
  c3=:SCENE=0,ALG=bpt,ITERS=10 \
  c4=:SCENE=3,ALG=bpm,W=2048,ITERS=4 \
  c5=:W=4096,MAXLEN=20,ITERS=2,FIRST=1 \
  > gpurun_out/c5_variants.log 2>&1; cat gpurun_out/c5_variants.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; tail -c 3000 gpurun_out/c5_bench.json; tail -3 gpurun_out/c5_bench.err

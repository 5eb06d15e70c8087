#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 90 --timeout-method=thread -p no:cacheprovider"
echo "== T1 group kernel"; timeout 300 $PYT tests/test_gpu_kernels.py -k "hashgrid and 7" > gpurun_out/c7_t1.log 2>&1; tail -6 gpurun_out/c7_t1.log
echo "== variants"
timeout 400 python tools/gpu_variants.py \
  lane=:SVCM_MERGE_KERNEL=1 \
  group=:SVCM_MERGE_KERNEL=7 \
  group_late=:SVCM_MERGE_KERNEL=7,FIRST=20,ITERS=10 \
  lane_late=:SVCM_MERGE_KERNEL=1,FIRST=20,ITERS=10 \
  c2_group=:SVCM_MERGE_KERNEL=7,W=512,ITERS=40 \
  c4_group=:SVCM_MERGE_KERNEL=7,SCENE=3,ALG=bpm,W=2048,ITERS=4 \
  c5_group=:SVCM_MERGE_KERNEL=7,W=4096,MAXLEN=20,ITERS=2,FIRST=1 \
  > gpurun_out/c7_variants.log 2>&1; cat gpurun_out/c7_variants.log
echo "== T2"; timeout 300 $PYT tests/test_gpu_render.py -k "merge_group or full_size_same" > gpurun_out/c7_t2.log 2>&1; tail -4 gpurun_out/c7_t2.log
echo "== ncu group"; SVCM_MERGE_KERNEL=7 timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_mergeGroup --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_merge_group python tools/run_iters.py > gpurun_out/c7_ncu1.log 2>&1; tail -2 gpurun_out/c7_ncu1.log

#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 90 --timeout-method=thread -p no:cacheprovider"
echo "== T1+T2 quick"; timeout 400 $PYT tests/test_gpu_kernels.py tests/test_gpu_render.py -k "(hashgrid and not cell) or merge_group or full_size_same" > gpurun_out/c6_t.log 2>&1; tail -4 gpurun_out/c6_t.log
echo "== variants"
timeout 500 python tools/gpu_variants.py \
  lane=:SVCM_MERGE_KERNEL=1 \
  lanev2=:SVCM_MERGE_KERNEL=6 \
  lanev2_bucketorder=:SVCM_MERGE_KERNEL=6,SVCM_QUERY_ORDER=bucket \
  isolated=:SVCM_MERGE_KERNEL=6,SVCM_NO_GRID_OVERLAP=1,SVCM_NO_PASS_OVERLAP=1 \
  isolated_legacybuild=:SVCM_MERGE_KERNEL=6,SVCM_NO_GRID_OVERLAP=1,SVCM_NO_PASS_OVERLAP=1,SVCM_LEGACY_BUILD=1 \
  c4_v2=:SVCM_MERGE_KERNEL=6,SCENE=3,ALG=bpm,W=2048,ITERS=4 \
  c5_v2=:SVCM_MERGE_KERNEL=6,W=4096,MAXLEN=20,ITERS=2,FIRST=1 \
  > gpurun_out/c6_variants.log 2>&1; cat gpurun_out/c6_variants.log
echo "== ncu lanev2"; SVCM_MERGE_KERNEL=6 timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_mergeLaneV2 --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_merge_lanev2 python tools/run_iters.py > gpurun_out/c6_ncu1.log 2>&1; tail -2 gpurun_out/c6_ncu1.log

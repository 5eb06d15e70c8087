#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 90 --timeout-method=thread -p no:cacheprovider"
echo "== T1 merge kernels"; timeout 300 $PYT tests/test_gpu_kernels.py -k "hashgrid and not cell" > gpurun_out/c4_t1.log 2>&1; tail -4 gpurun_out/c4_t1.log
echo "== variants"
timeout 600 python tools/gpu_variants.py \
  lane_spatial=:SVCM_MERGE_KERNEL=1 \
  lane_bucketorder=:SVCM_MERGE_KERNEL=1,SVCM_QUERY_ORDER=bucket \
  staged512=:SVCM_MERGE_KERNEL=2 \
  staged512_bucketorder=:SVCM_MERGE_KERNEL=2,SVCM_QUERY_ORDER=bucket \
  staged256=:SVCM_MERGE_KERNEL=5 \
  staged128=:SVCM_MERGE_KERNEL=4 \
  staged_plain=:SVCM_MERGE_KERNEL=3 \
  c4_lane=:SVCM_MERGE_KERNEL=1,SCENE=3,ALG=bpm,W=2048,ITERS=4 \
  c4_staged=:SVCM_MERGE_KERNEL=2,SCENE=3,ALG=bpm,W=2048,ITERS=4 \
  c5_lane=:SVCM_MERGE_KERNEL=1,W=4096,MAXLEN=20,ITERS=2,FIRST=1 \
  c5_staged=:SVCM_MERGE_KERNEL=2,W=4096,MAXLEN=20,ITERS=2,FIRST=1 \
  lazy=build/variants/libsvcm_merge-lazy-ranges.so:SVCM_ALLOW_OLD_LIBRARY=1 \
  r1=build/variants/libsvcm_r1.so:SVCM_ALLOW_OLD_LIBRARY=1 \
  camocc=build/variants/libsvcm_camera-occupancy.so:SVCM_ALLOW_OLD_LIBRARY=1 \
  > gpurun_out/c4_variants.log 2>&1; cat gpurun_out/c4_variants.log
echo "== ncu staged"; SVCM_MERGE_KERNEL=2 timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_mergeBucketStaged --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_merge_staged_b python tools/run_iters.py > gpurun_out/c4_ncu1.log 2>&1; tail -2 gpurun_out/c4_ncu1.log
echo "== T2"; timeout 400 $PYT tests/test_gpu_render.py -k "t2_same_stream or merge_group or full_size_same or path_length" > gpurun_out/c4_t2.log 2>&1; tail -6 gpurun_out/c4_t2.log

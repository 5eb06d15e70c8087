#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 90 --timeout-method=thread -p no:cacheprovider"
echo "== ncu staged"; SVCM_MERGE_KERNEL=2 timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_mergeBucketStaged --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_merge_staged python tools/run_iters.py > gpurun_out/c3_ncu1.log 2>&1; tail -2 gpurun_out/c3_ncu1.log
echo "== ncu lane"; SVCM_MERGE_KERNEL=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_mergeLane --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_merge_lane python tools/run_iters.py > gpurun_out/c3_ncu2.log 2>&1; tail -2 gpurun_out/c3_ncu2.log
echo "== variants"
timeout 300 python tools/gpu_variants.py \
  newbuild=:SVCM_MERGE_KERNEL=1 \
  legacybuild=:SVCM_MERGE_KERNEL=1,SVCM_LEGACY_BUILD=1 \
  newbuild_serial=:SVCM_MERGE_KERNEL=1,SVCM_NO_PASS_OVERLAP=1 \
  > gpurun_out/c3_variants.log 2>&1; cat gpurun_out/c3_variants.log
echo "== tests"; timeout 300 $PYT tests/test_gpu_render.py tests/test_gpu_kernels.py -k "full_size_same or hashgrid_build or grid_of_full" > gpurun_out/c3_tests.log 2>&1; tail -5 gpurun_out/c3_tests.log

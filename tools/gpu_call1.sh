#!/bin/bash
# first GPU call of round 2: new cell-layout tests, A/B of variants, sanitizer on a tiny case, full suite
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi1.log 2>&1
echo "== new tests"; timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_render.py -q -m gpu --timeout 300 -k "hashgrid or cell_layout" > gpurun_out/t_new.log 2>&1; tail -15 gpurun_out/t_new.log
echo "== variants"
rm -f gpurun_out/variants.jsonl
timeout 900 python tools/gpu_variants.py \
  r1=build/variants/libsvcm_r1.so: \
  lazy=build/variants/libsvcm_merge-lazy-ranges.so: \
  gridgather=build/variants/libsvcm_grid-gather.so: \
  camocc=build/variants/libsvcm_camera-occupancy.so: \
  bucket=:SVCM_GRID_LAYOUT=bucket \
  cell0=:SVCM_GRID_LAYOUT=cell,SVCM_MERGE_VARIANT=0 \
  cell1=:SVCM_GRID_LAYOUT=cell,SVCM_MERGE_VARIANT=1 \
  cell2=:SVCM_GRID_LAYOUT=cell,SVCM_MERGE_VARIANT=2 \
  cell3=:SVCM_GRID_LAYOUT=cell,SVCM_MERGE_VARIANT=3 \
  c4_bucket=:SVCM_GRID_LAYOUT=bucket,SCENE=3,ALG=bpm,W=2048,ITERS=4 \
  c4_cell1=:SVCM_GRID_LAYOUT=cell,SVCM_MERGE_VARIANT=1,SCENE=3,ALG=bpm,W=2048,ITERS=4 \
  > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log
echo "== sanitizer"
SVCM_GRID_LAYOUT=cell timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "collisions and 64 and cell" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/memcheck.log
echo "== full suite (defaults)"; timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/t_all.log 2>&1; tail -8 gpurun_out/t_all.log

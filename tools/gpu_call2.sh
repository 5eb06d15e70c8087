#!/bin/bash
# single-GPU check of the cell layout, the stream overlap and the bench line; every step bounded
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 120 --timeout-method=thread -p no:cacheprovider"
echo "== T1 cell"; timeout 420 $PYT tests/test_gpu_kernels.py -k "hashgrid or lights" > gpurun_out/c2_t1.log 2>&1; tail -6 gpurun_out/c2_t1.log
echo "== T2 cell"; timeout 420 $PYT tests/test_gpu_render.py -k "cell_layout or t2_same_stream" > gpurun_out/c2_t2.log 2>&1; tail -6 gpurun_out/c2_t2.log
echo "== variants"
rm -f gpurun_out/variants.jsonl
timeout 600 python tools/gpu_variants.py \
  r1=build/variants/libsvcm_r1.so: \
  camocc=build/variants/libsvcm_camera-occupancy.so: \
  bucket=:SVCM_GRID_LAYOUT=bucket \
  bucket_serial=:SVCM_GRID_LAYOUT=bucket,SVCM_NO_PASS_OVERLAP=1 \
  cell0=:SVCM_GRID_LAYOUT=cell,SVCM_MERGE_VARIANT=0 \
  cell1=:SVCM_GRID_LAYOUT=cell,SVCM_MERGE_VARIANT=1 \
  cell2=:SVCM_GRID_LAYOUT=cell,SVCM_MERGE_VARIANT=2 \
  cell3=:SVCM_GRID_LAYOUT=cell,SVCM_MERGE_VARIANT=3 \
  cell1_serial=:SVCM_GRID_LAYOUT=cell,SVCM_MERGE_VARIANT=1,SVCM_NO_PASS_OVERLAP=1 \
  c4_bucket=:SVCM_GRID_LAYOUT=bucket,SCENE=3,ALG=bpm,W=2048,ITERS=4 \
  c4_cell1=:SVCM_GRID_LAYOUT=cell,SVCM_MERGE_VARIANT=1,SCENE=3,ALG=bpm,W=2048,ITERS=4 \
  c5_bucket=:SVCM_GRID_LAYOUT=bucket,W=4096,MAXLEN=20,ITERS=2,FIRST=1 \
  c5_cell1=:SVCM_GRID_LAYOUT=cell,SVCM_MERGE_VARIANT=1,W=4096,MAXLEN=20,ITERS=2,FIRST=1 \
  > gpurun_out/c2_variants.log 2>&1; cat gpurun_out/c2_variants.log
echo "== full suite, cell layout"; SVCM_GRID_LAYOUT=cell timeout 600 $PYT tests > gpurun_out/c2_all_cell.log 2>&1; tail -8 gpurun_out/c2_all_cell.log
echo "== full suite, defaults"; timeout 600 $PYT tests > gpurun_out/c2_all.log 2>&1; tail -8 gpurun_out/c2_all.log

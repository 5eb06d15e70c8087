#!/bin/bash
# single-GPU check: staged bucket merge, stream overlap, cell layout, bench line; every step bounded
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 90 --timeout-method=thread -p no:cacheprovider"
echo "== T1 merge kernels"; timeout 300 $PYT tests/test_gpu_kernels.py -k "hashgrid and not cell" > gpurun_out/c2_t1.log 2>&1; tail -4 gpurun_out/c2_t1.log
echo "== variants"
rm -f gpurun_out/variants.jsonl
timeout 500 python tools/gpu_variants.py \
  lane=:SVCM_MERGE_KERNEL=1 \
  lane_serial=:SVCM_MERGE_KERNEL=1,SVCM_NO_PASS_OVERLAP=1 \
  staged=:SVCM_MERGE_KERNEL=2 \
  staged_plain=:SVCM_MERGE_KERNEL=3 \
  staged_small=:SVCM_MERGE_KERNEL=4 \
  c4_lane=:SVCM_MERGE_KERNEL=1,SCENE=3,ALG=bpm,W=2048,ITERS=4 \
  c4_staged=:SVCM_MERGE_KERNEL=2,SCENE=3,ALG=bpm,W=2048,ITERS=4 \
  c5_lane=:SVCM_MERGE_KERNEL=1,W=4096,MAXLEN=20,ITERS=2,FIRST=1 \
  c5_staged=:SVCM_MERGE_KERNEL=2,W=4096,MAXLEN=20,ITERS=2,FIRST=1 \
  c3_bpt=:SCENE=0,ALG=bpt,ITERS=10 \
  c3_bpt_serial=:SCENE=0,ALG=bpt,ITERS=10,SVCM_NO_PASS_OVERLAP=1 \
  > gpurun_out/c2_variants.log 2>&1; cat gpurun_out/c2_variants.log
echo "== T1 rest + T2"; timeout 400 $PYT tests/test_gpu_kernels.py tests/test_gpu_render.py -k "not cell and not hashgrid and not full_size and not drop_in" > gpurun_out/c2_t2.log 2>&1; tail -6 gpurun_out/c2_t2.log
echo "== full size"; timeout 400 $PYT tests/test_gpu_render.py -k "full_size or drop_in" > gpurun_out/c2_full.log 2>&1; tail -6 gpurun_out/c2_full.log
echo "== other files"; timeout 300 $PYT tests/test_bvh.py tests/test_image.py tests/test_abi.py > gpurun_out/c2_other.log 2>&1; tail -4 gpurun_out/c2_other.log
echo "== cell layout"; timeout 300 $PYT tests/test_gpu_kernels.py tests/test_gpu_render.py -k "cell" > gpurun_out/c2_cell.log 2>&1; tail -4 gpurun_out/c2_cell.log

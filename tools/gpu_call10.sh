#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 90 --timeout-method=thread -p no:cacheprovider"
echo "== tests"; timeout 500 $PYT tests/test_gpu_kernels.py tests/test_gpu_render.py -k "not cell and not drop_in and not 4096" > gpurun_out/c10_t.log 2>&1; tail -5 gpurun_out/c10_t.log
echo "== variants"
timeout 300 python tools/gpu_variants.py \
  default=: \
  isolated=:SVCM_NO_GRID_OVERLAP=1,SVCM_NO_PASS_OVERLAP=1 \
  c5=:W=4096,MAXLEN=20,ITERS=2,FIRST=1 \
  > gpurun_out/c10_variants.log 2>&1; cat gpurun_out/c10_variants.log

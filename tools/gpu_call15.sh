#!/bin/bash
# source-level profiles of the bounce kernels (second bounce of iteration 3) and of k_connect
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export ITERS=4
echo "== camera bounce 2"; timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_cameraBounce --launch-skip 31 --launch-count 1 -f -o gpurun_out/r2_camera_bounce2 python tools/run_iters.py > gpurun_out/c15_ncu1.log 2>&1; tail -1 gpurun_out/c15_ncu1.log
echo "== light bounce 2"; timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_lightBounce --launch-skip 28 --launch-count 1 -f -o gpurun_out/r2_light_bounce2 python tools/run_iters.py > gpurun_out/c15_ncu2.log 2>&1; tail -1 gpurun_out/c15_ncu2.log
echo "== connect"; timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_connect --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_connect python tools/run_iters.py > gpurun_out/c15_ncu3.log 2>&1; tail -1 gpurun_out/c15_ncu3.log
echo "== shadow (camera pass)"; timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_shadow --launch-skip 7 --launch-count 1 -f -o gpurun_out/r2_shadow python tools/run_iters.py > gpurun_out/c15_ncu4.log 2>&1; tail -1 gpurun_out/c15_ncu4.log
ls -la gpurun_out/*.ncu-rep

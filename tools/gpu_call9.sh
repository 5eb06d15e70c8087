#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== launch list"; ITERS=5 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python tools/run_iters.py > gpurun_out/c9_ncu0.log 2>&1; tail -2 gpurun_out/c9_ncu0.log
echo "== ncu lane spatial"; timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_mergeLane --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_merge_lane_spatial python tools/run_iters.py > gpurun_out/c9_ncu1.log 2>&1; tail -2 gpurun_out/c9_ncu1.log
echo "== ncu one iteration, all kernels"; ITERS=4 timeout 600 ncu --set full --clock-control none --launch-skip 96 --launch-count 32 -f -o gpurun_out/r2_iteration python tools/run_iters.py > gpurun_out/c9_ncu2.log 2>&1; tail -2 gpurun_out/c9_ncu2.log

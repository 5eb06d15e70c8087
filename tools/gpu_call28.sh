#!/bin/bash
# last commit: the full GPU suite and the launch list of the bench command
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 180 --timeout-method=thread -p no:cacheprovider"
echo "== full suite"; timeout 600 $PYT tests > gpurun_out/c28_tests.log 2>&1; tail -3 gpurun_out/c28_tests.log
echo "== launch list of the bench command"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c28_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/c28_ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/c28_launches.csv 1 > gpurun_out/c28_launch_shares.txt 2>/dev/null; tail -15 gpurun_out/c28_launch_shares.txt

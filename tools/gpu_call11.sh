#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 120 --timeout-method=thread -p no:cacheprovider"
echo "== bench"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err; tail -c 1800 gpurun_out/c11_bench.json; tail -3 gpurun_out/c11_bench.err
echo "== full suite"; timeout 700 $PYT tests > gpurun_out/c11_all.log 2>&1; tail -6 gpurun_out/c11_all.log
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

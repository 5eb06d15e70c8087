#!/bin/bash
# closing numbers after the grid-size change: a test subset, then the bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 180 --timeout-method=thread -p no:cacheprovider"
echo "== tests"; timeout 400 $PYT tests/test_gpu_kernels.py tests/test_gpu_render.py -k "intersect or runs or t2_same_stream or stream_schedules or hashgrid_query or shadow_queue" > gpurun_out/c27_tests.log 2>&1; tail -3 gpurun_out/c27_tests.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/c27_bench.json 2> gpurun_out/c27_bench.err; tail -c 300 gpurun_out/c27_bench.json; tail -2 gpurun_out/c27_bench.err

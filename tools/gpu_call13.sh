#!/bin/bash
# ray queries: plane-distance early-out + skippable primitive runs — exactness against the previous build, tests, timings
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 120 --timeout-method=thread -p no:cacheprovider"
echo "== A/B ray queries (previous build vs this one)"; timeout 200 python tools/ab_intersect.py build/variants/libsvcm_r2base.so smallvcm_b200/lib/libsvcm.so 2>&1 | tail -3
echo "== kernel + render tests"; timeout 600 $PYT tests/test_gpu_kernels.py tests/test_gpu_render.py -k "intersect or runs or bsdf or lights or t2_same_stream or t3 or path_length or merge_group" > gpurun_out/c13_tests.log 2>&1; tail -5 gpurun_out/c13_tests.log
echo "== variants"; rm -f gpurun_out/variants.jsonl
timeout 400 python tools/gpu_variants.py "r2base=build/variants/libsvcm_r2base.so:SVCM_ALLOW_OLD_LIBRARY=1" "new=" "new_noruns=:SVCM_DEBUG_NO_RUNS=1" "new_isolated=:OPTS=7=0" "r2base_isolated=build/variants/libsvcm_r2base.so:SVCM_ALLOW_OLD_LIBRARY=1,OPTS=7=0" 2>&1 | cut -c1-700
cp gpurun_out/variants.jsonl gpurun_out/c13_variants.jsonl
echo "== launch list (2 iterations)"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/c13_launches.csv env ITERS=2 python tools/run_iters.py > gpurun_out/c13_ncu.log 2>&1; python tools/launch_shares.py gpurun_out/c13_launches.csv 2 2>/dev/null | head -30

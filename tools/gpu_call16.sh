#!/bin/bash
# two-phase closest hit (per-lane candidate masks, staged primitives): exactness against the round's first build, tests, timings
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 120 --timeout-method=thread -p no:cacheprovider"
echo "== A/B ray queries (build before the ray-query changes vs this one)"; timeout 200 python tools/ab_intersect.py build/variants/libsvcm_r2base.so smallvcm_b200/lib/libsvcm.so 2>&1 | tail -3
echo "== kernel + render tests"; timeout 600 $PYT tests/test_gpu_kernels.py tests/test_gpu_render.py tests/test_bvh.py -k "intersect or runs or bsdf or lights or t2_same_stream or t3 or path_length or merge_group or bvh or traversal or hierarchy" > gpurun_out/c16_tests.log 2>&1; tail -5 gpurun_out/c16_tests.log
echo "== variants"; rm -f gpurun_out/variants.jsonl
timeout 400 python tools/gpu_variants.py "prev_iso=build/variants/libsvcm_prev.so:OPTS=7=0" "new_iso=:OPTS=7=0" "new_plain_iso=:OPTS=7=0,SVCM_DEBUG_NO_RUNS=1" "prev=build/variants/libsvcm_prev.so:" "new=" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k in ('name','ms_per_iter','light','build','camera','merge','error','mean','lightVertices')})
"
cp gpurun_out/variants.jsonl gpurun_out/c16_variants.jsonl
echo "== launch list (2 iterations)"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/c16_launches.csv env ITERS=2 python tools/run_iters.py > gpurun_out/c16_ncu.log 2>&1; python tools/launch_shares.py gpurun_out/c16_launches.csv 2 2>/dev/null | head -30

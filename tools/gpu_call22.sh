#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 180 --timeout-method=thread -p no:cacheprovider"
echo "== schedules (3x)"; for k in 1 2 3; do timeout 300 $PYT tests/test_gpu_render.py -k "stream_schedules or merge_group or multi_handle or cell_layout or accumulate" 2>&1 | tail -1; done
echo "== full suite"; timeout 900 $PYT tests > gpurun_out/c22_all.log 2>&1; tail -4 gpurun_out/c22_all.log
echo "== variants"; rm -f gpurun_out/variants.jsonl
timeout 400 python tools/gpu_variants.py "mask7=:OPTS=7=7" "mask15_single=:SVCM_SINGLE_QUERY_BUFFER=1" "mask15=" "mask15_512=:W=512" "mask15_bpm=:SCENE=3,ALG=bpm,W=2048" "mask15_256=:W=256,ITERS=60" "mask7_256=:W=256,ITERS=60,OPTS=7=7" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k in ('name','ms_per_iter','mpaths_s','light','build','camera','merge','error','mean','lightVertices')})
"
cp gpurun_out/variants.jsonl gpurun_out/c22_variants.jsonl

#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PYT="python -m pytest -q -m gpu --timeout 180 --timeout-method=thread -p no:cacheprovider"
echo "== full suite"; timeout 900 $PYT tests > gpurun_out/c18_all.log 2>&1; tail -6 gpurun_out/c18_all.log
echo "== schedules again (3x: races do not always show)"; for k in 1 2 3; do timeout 300 $PYT tests/test_gpu_render.py -k "stream_schedules or merge_group or multi_handle or cell_layout" 2>&1 | tail -1; done
echo "== variants"; rm -f gpurun_out/variants.jsonl
timeout 400 python tools/gpu_variants.py "mask7=:OPTS=7=7" "mask15=" "mask7_bpm=:OPTS=7=7,SCENE=3,ALG=bpm,W=2048" "mask15_bpm=:SCENE=3,ALG=bpm,W=2048" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k in ('name','ms_per_iter','mpaths_s','light','build','camera','merge','error','mean','lightVertices')})
"
cp gpurun_out/variants.jsonl gpurun_out/c18_variants.jsonl

#!/bin/bash
# grid sizes of the hot kernels (blocks per SM of the persistent grids): stage times alone and in the schedule
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/variants.jsonl
ISO="OPTS=7=0"
timeout 900 python tools/gpu_variants.py "base_iso=:$ISO" \
  "A_iso=:$ISO,SVCM_GRID_LIGHT=3,SVCM_GRID_CAMERA=3,SVCM_GRID_MERGE=8" \
  "B_iso=:$ISO,SVCM_GRID_LIGHT=6,SVCM_GRID_CAMERA=6,SVCM_GRID_MERGE=12" \
  "C_iso=:$ISO,SVCM_GRID_LIGHT=12,SVCM_GRID_CAMERA=12,SVCM_GRID_MERGE=16" \
  "D_iso=:$ISO,SVCM_GRID_LIGHT=16,SVCM_GRID_CAMERA=16" \
  "E_iso=:$ISO,SVCM_GRID_LIGHT=0,SVCM_GRID_CAMERA=0" \
  "N4_iso=:$ISO,SVCM_GRID_CONNECT=4" "N12_iso=:$ISO,SVCM_GRID_CONNECT=12" "N16_iso=:$ISO,SVCM_GRID_CONNECT=16" \
  "S5_iso=:$ISO,SVCM_GRID_SHADOW=5" "S10_iso=:$ISO,SVCM_GRID_SHADOW=10" "S16_iso=:$ISO,SVCM_GRID_SHADOW=16" \
  "base=" "A=:SVCM_GRID_LIGHT=3,SVCM_GRID_CAMERA=3" "B=:SVCM_GRID_LIGHT=6,SVCM_GRID_CAMERA=6" "C=:SVCM_GRID_LIGHT=12,SVCM_GRID_CAMERA=12" \
  "E=:SVCM_GRID_LIGHT=0,SVCM_GRID_CAMERA=0" "NS=:SVCM_GRID_CONNECT=16,SVCM_GRID_SHADOW=16" "base2=" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k in ('name','ms_per_iter','light','build','camera','merge','error')})
"
cp gpurun_out/variants.jsonl gpurun_out/c26_variants.jsonl

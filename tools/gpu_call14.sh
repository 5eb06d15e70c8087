#!/bin/bash
# occupancy A/B of the hot kernels (launch bounds variants), stage times alone and in the schedule
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/variants.jsonl
V=""
for n in l2 l4 c2 c4 m5 m6 n4; do V="$V ${n}_iso=build/variants/libsvcm_$n.so:OPTS=7=0"; done
timeout 900 python tools/gpu_variants.py "base_iso=:OPTS=7=0" $V "base=" "l2=build/variants/libsvcm_l2.so:" "m5=build/variants/libsvcm_m5.so:" "c4=build/variants/libsvcm_c4.so:" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k in ('name','ms_per_iter','light','build','camera','merge','error')})
"
cp gpurun_out/variants.jsonl gpurun_out/c14_variants.jsonl

#!/bin/bash
# the round's closing single-GPU run: full GPU suite, smoke, both bench arms, launch list of the bench command,
# one ncu --set full pass over every kernel of one iteration (summarised on the box, the report stays there)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=${TAG:-final}
PYT="python -m pytest -q -m gpu --timeout 180 --timeout-method=thread -p no:cacheprovider"
echo "== full suite"; timeout 900 $PYT tests > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (reference arm)"; SVCM_REF_BUDGET_S=40 timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 600 gpurun_out/${T}_bench_reference.json
echo "== bench"; timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 1500 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
echo "== launch list of the bench command"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/${T}_ncu_bench.log 2>&1; tail -1 gpurun_out/${T}_ncu_bench.log | cut -c1-200
python tools/launch_shares.py gpurun_out/${T}_launches.csv 1 > gpurun_out/${T}_launch_shares.txt 2>/dev/null; tail -16 gpurun_out/${T}_launch_shares.txt
echo "== ncu, every kernel of iteration 3"; ITERS=4 timeout 600 ncu --set full --clock-control none --launch-skip 96 --launch-count 32 -f -o /tmp/${T}_iteration python tools/run_iters.py > gpurun_out/${T}_ncu_iter.log 2>&1; tail -1 gpurun_out/${T}_ncu_iter.log
python tools/ncu_summary.py /tmp/${T}_iteration.ncu-rep > gpurun_out/${T}_iteration_kernel_summary.txt 2>&1; wc -l gpurun_out/${T}_iteration_kernel_summary.txt

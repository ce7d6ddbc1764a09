#!/bin/bash
# round 2, GPU call 1: all GPU tests, smoke, first bench lines of the new selection / exchange path, launch list of cfg1
set -u
OUT=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $OUT/r02_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/r02_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/r02_smoke.log
WORKLOADS="cfg1 cfg2 cfg4" bash profiles/bench_all.sh r02a
WORKLOADS="train8 train64" bash profiles/bench_all.sh r02a --no-e2e
CMD="python bench.py --workload cfg1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $OUT/plain_cfg1_r02a.json 2> $OUT/plain_cfg1_r02a.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_cfg1_r02a.csv $CMD > $OUT/ncu_launches_cfg1_r02a.log 2>&1
echo "launch list exit $?"
python profiles/summarize.py launches $OUT/launches_cfg1_r02a.csv > $OUT/launches_cfg1_r02a.txt 2>&1; tail -15 $OUT/launches_cfg1_r02a.txt

#!/bin/bash
# round 2, GPU call 10 (1 GPU): GPU tests after the register-cached select, launch lists of cfg1 (one-call pass) and train8
set -u
OUT=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $OUT/r02_pytest_gpu_g.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/r02_pytest_gpu_g.log
python profiles/pass_breakdown.py > $OUT/r02_pass_breakdown_g.txt 2>&1; cat $OUT/r02_pass_breakdown_g.txt
for wl in cfg1 cfg4; do
CMD="python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $OUT/plain_${wl}_r02g.json 2> $OUT/plain_${wl}_r02g.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_${wl}_r02.csv $CMD > $OUT/ncu_launches_${wl}_r02.log 2>&1
echo "launch list $wl exit $?"
python profiles/summarize.py launches $OUT/launches_${wl}_r02.csv > $OUT/launches_${wl}_r02.txt 2>&1; cat $OUT/launches_${wl}_r02.txt
done

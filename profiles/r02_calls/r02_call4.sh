#!/bin/bash
# round 2, GPU call 4 (1 GPU): GPU tests after the streamed-MC fixes, pass breakdown, fused-head bring-up variants,
# full ncu capture of the training-path (maps) kernel on batches of 8
set -u
OUT=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $OUT/r02_pytest_gpu_c.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/r02_pytest_gpu_c.log
python profiles/pass_breakdown.py > $OUT/r02_pass_breakdown.txt 2>&1; cat $OUT/r02_pass_breakdown.txt
WORKLOADS="cfg1" bash profiles/bench_all.sh r02c --no-e2e
WORKLOADS="cfg2h cfg1h" bash profiles/bench_all.sh r02c --no-e2e
for tag in hA hB hC; do
  ALS_LIB_TAG=$tag timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-150
  ALS_LIB_TAG=$tag WORKLOADS="cfg2h cfg1h" bash profiles/bench_all.sh r02c_$tag --no-e2e
done
CMD="python bench.py --workload train8 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $OUT/plain_train8_r02c.json 2> $OUT/plain_train8_r02c.err &&
ncu --set full --clock-control none --import-source on -k regex:score_tiles -s 40 -c 1 -f -o $OUT/prof_train8_r02c $CMD > $OUT/ncu_full_train8_r02c.log 2>&1
echo "ncu exit $?"
python profiles/summarize.py full $OUT/prof_train8_r02c.ncu-rep > $OUT/ncu_full_train8_r02c.txt 2>&1
python profiles/stalls.py $OUT/prof_train8_r02c.ncu-rep 0 30 > $OUT/stalls_train8_r02c.txt 2>&1
rm -f $OUT/prof_train8_r02c.ncu-rep
head -30 $OUT/ncu_full_train8_r02c.txt

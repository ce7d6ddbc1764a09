set -u
OUT=gpurun_out
# 1. full benches with e2e for the fused-head workloads
WORKLOADS="cfg1h cfg3h cfg4h" bash profiles/bench_all.sh h9
# 2. launch list + full ncu capture of the head kernel
CMD="python bench.py --workload cfg1h --pool 850 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $OUT/plain_cfg1h_r01.json 2> $OUT/plain_cfg1h_r01.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_cfg1h_r01.csv $CMD > $OUT/ncu_launches_cfg1h_r01.log 2>&1
echo "launch list exit $?"
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_head -s 4 -c 1 -f -o $OUT/prof_cfg1h_r01 $CMD > $OUT/ncu_full_cfg1h_r01.log 2>&1
echo "full capture exit $?"

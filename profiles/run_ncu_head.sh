set -u
OUT=gpurun_out
CMD="python bench.py --workload cfg1h --pool 850 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $OUT/plain_cfg1h_a.json 2> $OUT/plain_cfg1h_a.err &&
ncu --set full --clock-control none --import-source on -k regex:score_head -s 4 -c 1 -f -o $OUT/prof_cfg1h_a $CMD > $OUT/ncu_full_cfg1h_a.log 2>&1
echo "full capture exit $?"
tail -3 $OUT/ncu_full_cfg1h_a.log

#!/bin/bash
set -u
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --workload cfg4 --dtype bf16"
$CMD > $OUT/plain_cfg4_bf16_r02.json 2> $OUT/plain_cfg4_bf16_r02.err &&
ncu --set full --clock-control none --import-source on -k regex:score_tiles -s 4 -c 1 -f -o $OUT/prof_cfg4_bf16_r02 $CMD > $OUT/ncu_full_cfg4_bf16_r02.log 2>&1
echo "capture exit $?"
python profiles/summarize.py full $OUT/prof_cfg4_bf16_r02.ncu-rep > $OUT/ncu_full_cfg4_bf16_r02.txt 2>&1
python profiles/stalls.py $OUT/prof_cfg4_bf16_r02.ncu-rep 0 45 > $OUT/stalls_cfg4_bf16_r02.txt 2>&1
rm -f $OUT/prof_cfg4_bf16_r02.ncu-rep
cat $OUT/ncu_full_cfg4_bf16_r02.txt | head -20; head -62 $OUT/stalls_cfg4_bf16_r02.txt

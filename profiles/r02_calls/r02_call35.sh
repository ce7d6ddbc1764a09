#!/bin/bash
# round 2, GPU call 35 (1 GPU): ncu captures after the guided-run tile scheduler -- cfg3 f32 (margin), cfg3 bf16, cfg4 bf16
set -u
OUT=gpurun_out
cap() {  # tag, kernel regex, skip, bench args...
  tag=$1; rx=$2; skip=$3; shift; shift; shift
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $*"
  $CMD > $OUT/plain_${tag}_r02.json 2> $OUT/plain_${tag}_r02.err &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $OUT/prof_${tag}_r02 $CMD > $OUT/ncu_full_${tag}_r02.log 2>&1
  echo "capture $tag exit $?"
  python profiles/summarize.py full $OUT/prof_${tag}_r02.ncu-rep > $OUT/ncu_full_${tag}_r02.txt 2>&1
  python profiles/stalls.py $OUT/prof_${tag}_r02.ncu-rep 0 25 > $OUT/stalls_${tag}_r02.txt 2>&1
  rm -f $OUT/prof_${tag}_r02.ncu-rep
}
cap cfg3_runs score_tiles 4 --workload cfg3
cap cfg3_bf16_runs score_tiles 4 --workload cfg3 --dtype bf16
cap cfg4_bf16_runs score_tiles 4 --workload cfg4 --dtype bf16
grep -E "time_duration|issue_active|pipe_xu|dram_throughput|dram__bytes_read.sum |warps_active" $OUT/ncu_full_cfg3_runs_r02.txt $OUT/ncu_full_cfg3_bf16_runs_r02.txt $OUT/ncu_full_cfg4_bf16_runs_r02.txt

#!/bin/bash
# round 2, GPU call 24 (1 GPU): final ncu captures -- fused head with T = 8 (cfg2h) and bf16 C = 6 after the F2I change (cfg4 bf16)
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
cap cfg2h score_head 3 --workload cfg2h --pool 170
cap cfg4_bf16_final score_tiles 4 --workload cfg4 --dtype bf16
grep -E "time_duration|issue_active|pipe_xu|pipe_tensor|dram_throughput|inst_executed.sum" $OUT/ncu_full_cfg2h_r02.txt $OUT/ncu_full_cfg4_bf16_final_r02.txt

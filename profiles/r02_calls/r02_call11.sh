#!/bin/bash
# round 2, GPU call 11 (1 GPU): guard-band memory-safety test, ncu captures: fused head T=1 (who makes the shared-memory
# "bank conflicts"), streamed MC update kernel, bf16 cfg1
set -u
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "outputs_stay or select" > $OUT/r02_pytest_guard.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r02_pytest_guard.log
cap() {  # tag, kernel regex, skip, bench args...
  tag=$1; rx=$2; skip=$3; shift; shift; shift
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $*"
  $CMD > $OUT/plain_${tag}_r02.json 2> $OUT/plain_${tag}_r02.err &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $OUT/prof_${tag}_r02 $CMD > $OUT/ncu_full_${tag}_r02.log 2>&1
  echo "capture $tag exit $?"
  python profiles/summarize.py full $OUT/prof_${tag}_r02.ncu-rep > $OUT/ncu_full_${tag}_r02.txt 2>&1
  python profiles/stalls.py $OUT/prof_${tag}_r02.ncu-rep 0 20 > $OUT/stalls_${tag}_r02.txt 2>&1
  python profiles/shared_access.py $OUT/prof_${tag}_r02.ncu-rep 0 > $OUT/shared_${tag}_r02.txt 2>&1
  rm -f $OUT/prof_${tag}_r02.ncu-rep
}
cap cfg1h score_head 4 --workload cfg1h --pool 850
cap cfg2s mc_update 12 --workload cfg2s --pool 350
cap cfg1_bf16 score_tiles 4 --workload cfg1 --dtype bf16
head -12 $OUT/shared_cfg1h_r02.txt | cut -c1-260
grep -E "time_duration|dram__bytes|dram_throughput|issue_active|pipe_xu|inst_executed.sum" $OUT/ncu_full_cfg2s_r02.txt $OUT/ncu_full_cfg1_bf16_r02.txt

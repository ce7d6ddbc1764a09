#!/bin/bash
# Profiling recipe (B200_PROFILING.md): launch list + one full capture of the dominant kernel.
# Usage (under gpurun): bash profiles/run_ncu.sh <workload> <tag> [extra bench args, e.g. --dtype bf16]
set -u
WL=${1:-cfg2}
TAG=${2:-r01}
shift; shift
EXTRA="$*"
OUT=gpurun_out
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA"
$CMD > $OUT/plain_${WL}_${TAG}.json 2> $OUT/plain_${WL}_${TAG}.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_${WL}_${TAG}.csv $CMD > $OUT/ncu_launches_${WL}_${TAG}.log 2>&1
echo "launch list exit $?"
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_tiles -s 4 -c 1 \
    -f -o $OUT/prof_${WL}_${TAG} $CMD > $OUT/ncu_full_${WL}_${TAG}.log 2>&1
echo "full capture exit $?"
# summaries are made on the box; the 20 MB report itself only travels back with KEEP_REP=1 (gpurun_out is capped at 64 MiB)
python profiles/summarize.py full $OUT/prof_${WL}_${TAG}.ncu-rep > $OUT/ncu_full_${WL}_${TAG}.txt 2>&1
python profiles/stalls.py $OUT/prof_${WL}_${TAG}.ncu-rep 0 20 > $OUT/stalls_${WL}_${TAG}.txt 2>&1
python profiles/summarize.py launches $OUT/launches_${WL}_${TAG}.csv > $OUT/launches_${WL}_${TAG}.txt 2>&1
[ "${KEEP_REP:-0}" = "1" ] || rm -f $OUT/prof_${WL}_${TAG}.ncu-rep
ls -la $OUT | tail -12

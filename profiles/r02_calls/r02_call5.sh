#!/bin/bash
# round 2, GPU call 5 (1 GPU): GPU tests, PDL ordering probe, one-call pool pass, fused-head bring-up variants (EPB=1, FMA-pipe exp2)
set -u
OUT=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $OUT/r02_pytest_gpu_d.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/r02_pytest_gpu_d.log
semanticsegmentationactivelearning_b200/build/pdl_probe > $OUT/r02_pdl_probe.txt 2>&1; cat $OUT/r02_pdl_probe.txt
python profiles/pass_breakdown.py > $OUT/r02_pass_breakdown_d.txt 2>&1; cat $OUT/r02_pass_breakdown_d.txt
WORKLOADS="cfg1 cfg4" bash profiles/bench_all.sh r02d --no-e2e
for tag in "" hD hE hP2 hP4; do
  [ -n "$tag" ] && { ALS_LIB_TAG=$tag timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-120; }
  for wl in cfg2h cfg1h; do
    ALS_LIB_TAG=$tag timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-e2e > $OUT/bench_${wl}_r02d_$tag.json 2> $OUT/bench_${wl}_r02d_$tag.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${wl}_r02d_$tag.json")); r=d["roofline"]
    print("$wl [$tag] value=%.2f Gpix/s launch_ms=%.3f sm=%s %s" % (d["value"], r["avg_launch_ms"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("$wl [$tag] no line", e)
PY
  done
done

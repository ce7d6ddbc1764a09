#!/bin/bash
# Run the five BASELINE workloads (1 GPU) and print one summary line each.  Usage: bash profiles/bench_all.sh <tag> [extra bench args]
TAG=${1:-run}; shift
for wl in ${WORKLOADS:-cfg1 cfg2 cfg3 cfg4 cfg5}; do
  timeout 400 python bench.py --workload $wl --no-cpu-baseline "$@" > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err || { echo "bench $wl FAILED"; tail -3 gpurun_out/bench_${wl}_${TAG}.err; continue; }
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${wl}_${TAG}.json"))
r=d["roofline"]; k=d["config"]["kernel"]
print("%s %s value=%.2f Gpix/s ms/step=%.3f kernel=%.0f GB/s frac=%.3f launch_ms=%.3f share=%.3f sm=%s MHz %s burst=%.0f GB/s e2e=%.3f grid=%d stages=%d smem=%d" % (
  "${wl}", d["dtype"], d["value"], d["ms_per_step"], r["achieved"], r["frac"], r["avg_launch_ms"], r["kernel_share_of_step"],
  d["clocks"]["sm_mhz"], d["clocks"]["reasons"], r["kernel_burst"]["GBps"], (d["e2e"] or {}).get("value", 0), k["grid"], k["stages"], k["smem_bytes"] or 0))
PY
done

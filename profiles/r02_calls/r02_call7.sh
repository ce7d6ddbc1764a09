#!/bin/bash
# round 2, GPU call 7 (1 GPU): after the stage-release fix -- determinism stress, GPU tests, perf of the hot kernels
set -u
OUT=gpurun_out
python profiles/determinism_stress.py 60 2>&1 | tee $OUT/r02_determinism_b.txt
python profiles/mc_localize.py 2>&1 | tee $OUT/r02_mc_localize_b.txt
(time timeout 1500 python -m pytest tests -m gpu -q) > $OUT/r02_pytest_gpu_f.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/r02_pytest_gpu_f.log
WORKLOADS="cfg1 cfg2 cfg3 cfg4 cfg5 cfg2s train8" bash profiles/bench_all.sh r02f --no-e2e
WORKLOADS="cfg1 cfg3 cfg4" bash profiles/bench_all.sh r02f_bf16 --no-e2e --dtype bf16
for wl in cfg2h cfg1h; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-e2e > $OUT/bench_${wl}_r02f.json 2> $OUT/bench_${wl}_r02f.err
  python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_${wl}_r02f.json") if l.startswith("{")][-1]); r=d["roofline"]
print("$wl value=%.2f Gpix/s launch_ms=%.3f sm=%s %s" % (d["value"], r["avg_launch_ms"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
PY
done

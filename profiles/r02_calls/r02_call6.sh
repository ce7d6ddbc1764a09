#!/bin/bash
# round 2, GPU call 6 (1 GPU): tests after the emit fast path, cfg1 one-call pass, train8/train64
set -u
OUT=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > $OUT/r02_pytest_gpu_e.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/r02_pytest_gpu_e.log
WORKLOADS="cfg1 train8 train64" bash profiles/bench_all.sh r02e --no-e2e
python bench.py --workload cfg1 --no-cpu-baseline --no-e2e --steps 20 > $OUT/bench_cfg1_r02e_s20.json 2>/dev/null; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_cfg1_r02e_s20.json") if l.startswith("{")][-1]); r=d["roofline"]
print("cfg1 20 steps: value=%.2f ms/step=%.4f launch_ms=%.4f share=%.3f ids=%s" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["kernel_share_of_step"], d["ids_check"]["ids_match_oracle"]))
PY
python profiles/h2d_ceiling.py > $OUT/r02_h2d_ceiling_g1.json 2>&1; cat $OUT/r02_h2d_ceiling_g1.json

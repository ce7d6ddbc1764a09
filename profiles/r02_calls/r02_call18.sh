#!/bin/bash
# round 2 final verification (1 GPU): smoke, every GPU test, default bench (both arms), training-path workloads
set -u
OUT=gpurun_out
timeout 300 python __graft_entry__.py --smoke > $OUT/r02_smoke_final.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/r02_smoke_final.log | cut -c1-200
(time timeout 900 python -m pytest tests -m gpu -q) > $OUT/r02_pytest_gpu_final.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/r02_pytest_gpu_final.log
python bench.py > $OUT/r02_bench_default.json 2> $OUT/r02_bench_default.err; echo "bench default exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02_bench_reference_arm.json 2> /dev/null; echo "reference arm exit $?"
WORKLOADS="train8 train64" bash profiles/bench_all.sh r02_final --no-e2e
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_default.json") if l.startswith("{")][-1]); r=d["roofline"]
print("default: value=%.3f ms=%.2f frac=%.3f share=%.4f e2e=%.4f cpu=%.5f launches=%s ids=%s" % (d["value"], d["ms_per_step"], r["frac"], r["kernel_share_of_step"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["gpu_launches"], d["ids_check"]["ids_match_oracle"]))
print("e2e_alt:", {k: round(v["value"], 4) for k, v in d["e2e_alt"].items()})
PY

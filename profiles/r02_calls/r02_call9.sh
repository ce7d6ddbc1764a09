#!/bin/bash
# round 2, GPU call 9 (1 GPU): the default bench line (both arms), launch list + full ncu capture of the dominant kernel (cfg2)
set -u
OUT=gpurun_out
python bench.py > $OUT/r02_bench_default.json 2> $OUT/r02_bench_default.err; echo "bench default exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02_bench_reference_arm.json 2> $OUT/r02_bench_reference_arm.err; echo "reference arm exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_default.json") if l.startswith("{")][-1]); r=d["roofline"]
print("default: value=%.3f ms=%.2f frac=%.3f share=%.4f e2e=%.4f cpu=%s launches=%s ids=%s" % (d["value"], d["ms_per_step"], r["frac"], r["kernel_share_of_step"], d["e2e"]["value"], d["cpu_baseline"], d["gpu_launches"], d["ids_check"]["ids_match_oracle"]))
print("e2e_alt:", {k: round(v["value"], 4) for k, v in d["e2e_alt"].items()})
d=json.loads([l for l in open("gpurun_out/r02_bench_reference_arm.json") if l.startswith("{")][-1])
print("reference arm: value=%.5f cores=%s" % (d["value"], d["cpu_baseline"]["cores"]))
PY
bash profiles/run_ncu.sh cfg2 r02

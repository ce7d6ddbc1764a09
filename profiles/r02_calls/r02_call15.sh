#!/bin/bash
set -u
OUT=gpurun_out
python profiles/host_staging.py 2>&1 | tee $OUT/r02_host_staging_b.txt
for n in 2 4 16; do echo "threads=$n"; ALS_STAGE_THREADS=$n python profiles/host_staging.py 2>&1 | tail -1; done
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r02_pytest_gpu_h.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r02_pytest_gpu_h.log
python bench.py --no-cpu-baseline > $OUT/r02_bench_default_b.json 2>/dev/null; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_default_b.json") if l.startswith("{")][-1])
print("e2e", round(d["e2e"]["value"],4), {k:(round(v["value"],4), round(v["h2d_GBps_per_gpu"],1)) for k,v in d["e2e_alt"].items()})
PY

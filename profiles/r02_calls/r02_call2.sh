#!/bin/bash
# round 2, GPU call 2 (2 GPUs): every GPU test incl. the torchrun world-2 check, then the 2-GPU bench lines
set -u
OUT=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $OUT/r02_pytest_gpu_g2.log 2>&1; echo "pytest exit $?"; tail -8 $OUT/r02_pytest_gpu_g2.log
for wl in cfg2 cfg3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_${wl}_g2_r02a.json 2> $OUT/bench_${wl}_g2_r02a.err
  echo "bench $wl g2 exit $?"; tail -c 600 $OUT/bench_${wl}_g2_r02a.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${wl}_g2_r02a.json")); r=d["roofline"]
    print("$wl g2 value=%.2f ms=%.3f share=%.4f frac=%.3f e2e=%s ids=%s" % (d["value"], d["ms_per_step"], r["kernel_share_of_step"], r["frac"], (d["e2e"] or {}).get("value"), d["ids_check"]))
except Exception as e: print("no line", e)
PY
done

#!/bin/bash
# round 2, GPU call 8 (8 GPUs): multi-GPU tests (single-process all-GPU + torchrun world 2 and 8), 8-GPU bench lines, H2D ceiling
set -u
OUT=gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q) > $OUT/r02_pytest_multi_g8.log 2>&1; echo "pytest multi exit $?"; tail -5 $OUT/r02_pytest_multi_g8.log
run() {  # workload, extra args
  wl=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 8 --workload $wl --steps 5 --warmup 3 --no-cpu-baseline "$@" > $OUT/bench_${wl}_g8_r02.json 2> $OUT/bench_${wl}_g8_r02.err
  echo "bench $wl g8 exit $?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_${wl}_g8_r02.json") if l.startswith("{")][-1]); r=d["roofline"]
    e=d["e2e"] or {}
    print("$wl g8 value=%.2f ms=%.3f share=%.4f frac=%.3f e2e=%s h2d/gpu=%s alt=%s ids=%s" % (d["value"], d["ms_per_step"], r["kernel_share_of_step"], r["frac"], e.get("value"), e.get("h2d_GBps_per_gpu"), {k:round(v["value"],3) for k,v in (d.get("e2e_alt") or {}).items()}, d["ids_check"]["ids_match_oracle"]))
except Exception as ex: print("no line", ex)
PY
}
run cfg2
run cfg3
run cfg5 --no-e2e
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 profiles/h2d_ceiling.py > $OUT/r02_h2d_ceiling_g8.json 2>$OUT/r02_h2d_ceiling_g8.err; tail -1 $OUT/r02_h2d_ceiling_g8.json

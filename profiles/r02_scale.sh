#!/bin/bash
# round 2 multi-GPU bench lines: bash profiles/r02_scale.sh <N> <workloads...>   (under gpurun --gpus N)
set -u
OUT=gpurun_out
N=$1; shift
for wl in "$@"; do
  extra=""; [ "$wl" = "cfg5" ] && extra="--no-e2e"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --workload $wl --steps 5 --warmup 3 --no-cpu-baseline $extra > $OUT/r02_bench_${wl}_g$N.json 2> $OUT/r02_bench_${wl}_g$N.err
  echo "bench $wl g$N exit $?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/r02_bench_${wl}_g$N.json") if l.startswith("{")][-1]); r=d["roofline"]; e=d["e2e"] or {}; pr=r["per_rank"]
    print("$wl g$N value=%.2f ms=%.3f share=%.4f frac=%.3f e2e=%s ids=%s" % (d["value"], d["ms_per_step"], r["kernel_share_of_step"], r["frac"], e.get("value"), d["ids_check"]["ids_match_oracle"]))
    print("   per rank scoring ms:", pr["scoring_ms_per_step"], "select call ms:", pr["select_call_ms"])
    print("   slowest rank scoring share of step: %.4f, exchange on slowest rank: %.3f ms" % (pr["slowest_rank_scoring_share_of_step"], pr["exchange_ms_on_slowest_rank"]))
except Exception as ex: print("no line", ex)
PY
done

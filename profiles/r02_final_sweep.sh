#!/bin/bash
# round 2 final sweep (1 GPU): every workload of bench.py with the shipped library -> gpurun_out/r02_bench_<workload>[_bf16].json
set -u
OUT=gpurun_out
one() {  # name, args...
  name=$1; shift
  timeout 500 python bench.py "$@" > $OUT/r02_bench_$name.json 2> $OUT/r02_bench_$name.err || { echo "$name FAILED"; tail -3 $OUT/r02_bench_$name.err; return; }
  python - <<PY
import json
d=json.loads([l for l in open("$OUT/r02_bench_$name.json") if l.startswith("{")][-1]); r=d["roofline"]; e=d.get("e2e") or {}
print("%-14s value=%8.3f Gpix/s ms/step=%9.3f kernel=%6.0f GB/s frac=%.3f share=%.4f e2e=%s alt=%s sm=%s %s ids=%s" % (
  "$name", d["value"], d["ms_per_step"], r["achieved"], r["frac"], r["kernel_share_of_step"], round(e["value"],4) if e else None,
  {k: round(v["value"],3) for k,v in (d.get("e2e_alt") or {}).items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
  (d.get("ids_check") or {}).get("ids_match_oracle")))
PY
}
for wl in cfg1 cfg2 cfg3 cfg4 cfg5; do one $wl --workload $wl --no-cpu-baseline; done
for wl in cfg1 cfg2 cfg3 cfg4 cfg5; do one ${wl}_bf16 --workload $wl --dtype bf16 --no-cpu-baseline --no-e2e; done
for wl in train8 train64 cfg2s; do one $wl --workload $wl --no-cpu-baseline --no-e2e; done
for wl in cfg1h cfg2h cfg3h cfg4h; do one $wl --workload $wl --no-cpu-baseline; done
one cfg1_cpu --workload cfg1 --no-e2e

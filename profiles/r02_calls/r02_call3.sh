#!/bin/bash
# round 2, GPU call 3 (1 GPU): full GPU test suite on the new ABI, fused finalize / zero-copy selection on cfg1 + train8,
# paired bf16 loads, FMA-pipe exp2 split (bring-up builds), streamed Monte-Carlo accumulation
set -u
OUT=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > $OUT/r02_pytest_gpu_b.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/r02_pytest_gpu_b.log
timeout 300 python __graft_entry__.py --smoke > $OUT/r02_smoke_b.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/r02_smoke_b.log
WORKLOADS="cfg1 cfg2s" bash profiles/bench_all.sh r02b --no-e2e
WORKLOADS="train8 train64" bash profiles/bench_all.sh r02b --no-e2e
WORKLOADS="cfg1 cfg3 cfg4" bash profiles/bench_all.sh r02b_bf16 --no-e2e --dtype bf16
for tag in poly2 poly4; do
  if [ -f semanticsegmentationactivelearning_b200/libalscore_$tag.so ]; then
    ALS_LIB_TAG=$tag WORKLOADS="cfg1 cfg3 cfg4" bash profiles/bench_all.sh r02b_bf16_$tag --no-e2e --dtype bf16
  fi
done

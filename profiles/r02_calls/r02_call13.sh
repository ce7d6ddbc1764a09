#!/bin/bash
# round 2, GPU call 13 (1 GPU): occupancy variants (pixels per thread cap / resident CTAs per SM) of the logits kernel
set -u
for tag in "" o24 o44 o25; do
  echo "== variant [$tag]"
  ALS_LIB_TAG=$tag WORKLOADS="cfg1 cfg4 cfg2" bash profiles/bench_all.sh r02i_$tag --no-e2e
  ALS_LIB_TAG=$tag WORKLOADS="cfg1 cfg4" bash profiles/bench_all.sh r02i_bf16_$tag --no-e2e --dtype bf16
done

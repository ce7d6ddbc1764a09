#!/bin/bash
set -u
for tag in "" mr40; do
  echo "== variant [$tag]"
  ALS_LIB_TAG=$tag WORKLOADS="cfg2 cfg5" bash profiles/bench_all.sh r02l_bf16_$tag --no-e2e --dtype bf16
done
ALS_LIB_TAG=mr40 timeout 300 python -m pytest tests/test_gpu_mc.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3

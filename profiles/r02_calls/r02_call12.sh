#!/bin/bash
# round 2, GPU call 12 (1 GPU): multi test (large-k exchange) at world 1, FMA-pipe exp2 for ONE class pair (bring-up build)
set -u
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $OUT/r02_pytest_multi_g1.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r02_pytest_multi_g1.log
WORKLOADS="cfg1 cfg3 cfg4" bash profiles/bench_all.sh r02h_bf16 --no-e2e --dtype bf16
ALS_LIB_TAG=poly1 WORKLOADS="cfg1 cfg3 cfg4" bash profiles/bench_all.sh r02h_bf16_poly1 --no-e2e --dtype bf16
ALS_LIB_TAG=poly1 WORKLOADS="cfg1 cfg4" bash profiles/bench_all.sh r02h_f32_poly1 --no-e2e

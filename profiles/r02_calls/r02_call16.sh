#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $OUT/r02_pytest_gpu_i.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r02_pytest_gpu_i.log
WORKLOADS="cfg1 cfg2 cfg3 cfg4 cfg5" bash profiles/bench_all.sh r02j_bf16 --no-e2e --dtype bf16
python profiles/determinism_stress.py 20 2>&1 | tail -3

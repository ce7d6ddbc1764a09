#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $OUT/r02_pytest_gpu_j.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r02_pytest_gpu_j.log
python profiles/emit_cost.py 2>&1 | tee $OUT/r02_emit_cost_b.txt
WORKLOADS="train8 train64 cfg1 cfg2" bash profiles/bench_all.sh r02k --no-e2e
python profiles/pass_breakdown.py 2>&1 | tail -3

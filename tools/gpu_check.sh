#!/bin/bash
# quick check of a build on one GPU box: GPU parity tests + the headline line
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --no-secondary --no-cpu-baseline > $O/bench_new.json 2> $O/bench_new.err

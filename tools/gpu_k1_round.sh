#!/bin/bash
# GPU parity tests, then the C4 shape (wide alphabet) and the headline, then ncu captures of the streaming kernels
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --workload c4 --no-secondary --no-cpu-baseline --steps 10 > $O/c4_main.json 2> $O/c4_main.err
timeout 300 python bench.py --no-secondary --no-cpu-baseline > $O/bench_new.json 2> $O/bench_new.err
B200CTC_NO_PIPELINE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2c_tokens_kernel -s 4 -c 1 -o $O/k1_wide -f \
  python bench.py --workload c4 --no-secondary --no-cpu-baseline --steps 2 --warmup 1 > $O/ncu_k1_wide.log 2>&1
ls -la $O/*.ncu-rep

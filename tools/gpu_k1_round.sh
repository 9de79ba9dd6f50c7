#!/bin/bash
# streaming-stage variants: GPU parity tests of the main build, then the C4 shape (wide alphabet) and the headline with each
# build under variants/ (register / occupancy variants of b2c_tokens_kernel), then ncu captures of both streaming kernels
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
V=$PWD/variants
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --workload c4 --no-secondary --no-cpu-baseline --steps 10 > $O/c4_main.json 2> $O/c4_main.err
for v in k1occ2 k1default; do
  if [ -f $V/libb200ctc_$v.so ]; then
    B200CTC_PROFILING_LIB=$V/libb200ctc_$v.so timeout 300 python bench.py --workload c4 --no-secondary --no-cpu-baseline --steps 10 > $O/c4_$v.json 2> $O/c4_$v.err
  fi
done
timeout 300 python bench.py --no-secondary --no-cpu-baseline > $O/bench_new.json 2> $O/bench_new.err
B200CTC_NO_PIPELINE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:b2c_tokens_tile -s 2 -c 1 -o $O/k1_tile -f \
  python bench.py --no-secondary --no-cpu-baseline --steps 2 --warmup 1 > $O/ncu_k1_tile.log 2>&1
B200CTC_NO_PIPELINE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2c_tokens_kernel -s 4 -c 1 -o $O/k1_wide -f \
  python bench.py --workload c4 --no-secondary --no-cpu-baseline --steps 2 --warmup 1 > $O/ncu_k1_wide.log 2>&1
ls -la $O/*.ncu-rep

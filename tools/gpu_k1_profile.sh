#!/bin/bash
# ncu captures of the streaming-stage kernels: lane-per-row tile kernel (C2, V = 32) and warp-per-row kernel (C4 shape, V = 1024)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B200CTC_NO_PIPELINE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:b2c_tokens_tile -s 2 -c 1 -o $O/k1_tile -f \
  python bench.py --no-secondary --no-cpu-baseline --steps 2 --warmup 1 > $O/ncu_k1_tile.log 2>&1
B200CTC_NO_PIPELINE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2c_tokens_kernel -s 4 -c 1 -o $O/k1_wide -f \
  python bench.py --workload c4 --no-secondary --no-cpu-baseline --steps 2 --warmup 1 > $O/ncu_k1_wide.log 2>&1
ls -la $O/*.ncu-rep

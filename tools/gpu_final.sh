#!/bin/bash
# End-of-round check on one GPU box, as the driver runs it: smoke(), the default bench (headline + secondary list) and
# the reference arm; everything lands in gpurun_out/.       gpurun --timeout 1500 -- 'bash tools/gpu_final.sh'
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
date
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
date
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" | tee -a $O/smoke.log
date
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time; echo "bench rc $?"
date
( time timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err ) 2> $O/bench_reference.time; echo "ref rc $?"
date
tail -n 4 $O/bench_default.time; tail -n 4 $O/bench_reference.time
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python bench.py --no-secondary --no-cpu-baseline --steps 2 --warmup 1 > $O/ncu_launches.log 2>&1
B200CTC_NO_PIPELINE=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:b2c_beam_fast -s 2 -c 1 -o $O/beam_full -f python bench.py --no-secondary --no-cpu-baseline --steps 2 --warmup 1 > $O/ncu_full.log 2>&1
date

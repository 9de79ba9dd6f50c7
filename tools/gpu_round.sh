#!/bin/bash
# One gpurun call of the round: GPU parity tests, headline bench with the variant libraries under variants/ (baseline of
# the previous commit, phase-clock build), selected secondary entries, ncu launch list and one full capture of the
# headline beam kernel.  Everything lands in gpurun_out/.       gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
V=$PWD/variants
echo "== pytest -m gpu"; date
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" | tee -a $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
echo "== headline, three builds"; date
if [ -f $V/libb200ctc_base.so ]; then
  B200CTC_PROFILING_LIB=$V/libb200ctc_base.so timeout 300 python bench.py --no-secondary --no-cpu-baseline > $O/bench_base.json 2> $O/bench_base.err
fi
B200CTC_NO_HINTED=1 timeout 300 python bench.py --no-secondary --no-cpu-baseline > $O/bench_new_nohint.json 2> $O/bench_new_nohint.err
timeout 300 python bench.py --no-secondary --no-cpu-baseline > $O/bench_new.json 2> $O/bench_new.err
B200CTC_HOST_PROFILE=1 timeout 300 python bench.py --no-secondary --no-cpu-baseline --steps 5 > /dev/null 2> $O/bench_new_hostprof.err
if [ -f $V/libb200ctc_clk.so ]; then
  B200CTC_PROFILING_LIB=$V/libb200ctc_clk.so timeout 300 python bench.py --no-secondary --no-cpu-baseline --steps 3 > $O/bench_clk.json 2> $O/phase_clocks.txt
fi
echo "== secondary entries"; date
timeout 600 python bench.py --no-cpu-baseline --secondary c3,beam10,beam50 > $O/bench_new_secondary.json 2> $O/bench_new_secondary.err
if [ -f $V/libb200ctc_base.so ]; then
  B200CTC_PROFILING_LIB=$V/libb200ctc_base.so timeout 600 python bench.py --no-cpu-baseline --secondary c4 > $O/bench_base_secondary.json 2> $O/bench_base_secondary.err
fi
echo "== ncu"; date
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $O/launches.csv \
  python bench.py --no-secondary --no-cpu-baseline --steps 2 --warmup 1 > $O/ncu_launches.log 2>&1
B200CTC_NO_PIPELINE=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:b2c_beam_fast -s 2 -c 1 -o $O/beam_full -f \
  python bench.py --no-secondary --no-cpu-baseline --steps 2 --warmup 1 > $O/ncu_full.log 2>&1
date
ls -la $O

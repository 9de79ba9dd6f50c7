#!/bin/bash
# End-of-round check on one GPU box, as the driver runs it: smoke(), the default bench (headline + secondary list) and
# the reference arm; everything lands in gpurun_out/.       gpurun --timeout 1500 -- 'bash tools/gpu_final.sh'
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
date
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" | tee -a $O/smoke.log
date
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time; echo "bench rc $?"
date
( time timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err ) 2> $O/bench_reference.time; echo "ref rc $?"
date
tail -3 $O/bench_default.time $O/bench_reference.time

#!/bin/bash
# Final evidence of a round in one short call: GPU parity suite, smoke(), the default bench line, the ncu launch list of the bench command.
TAG=${1:-final}
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log; tail -4 $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke exit $?"
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"; cat $O/bench_$TAG.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --slots 1 > $O/bench_under_ncu_$TAG.log 2>&1; echo "launch list exit $?"
python tools/launch_summary.py $O/launches_$TAG.csv > $O/launch_summary_$TAG.txt 2>&1; head -14 $O/launch_summary_$TAG.txt

#!/bin/bash
# round 2 final (second half): everything the driver runs at round end on the final tree, plus the launch list of one proof
TAG=${1:-r02final2}
O=gpurun_out; mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log; tail -6 $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 $O/smoke_$TAG.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"; cat $O/bench_$TAG.json | cut -c1-300; tail -3 $O/bench_$TAG.err
timeout 300 python tools/latency_probe.py > $O/latency_$TAG.jsonl 2>/dev/null; cat $O/latency_$TAG.jsonl
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-job-records --slots 1 > $O/bench_under_ncu_$TAG.log 2>&1; echo "launch list exit $?"
python tools/launch_summary.py $O/launches_$TAG.csv > $O/launch_summary_$TAG.txt 2>&1; head -16 $O/launch_summary_$TAG.txt

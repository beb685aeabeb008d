#!/bin/bash
TAG=${1:-r02g}
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests/test_gpu_prover.py tests/test_gpu_tasks.py tests/test_gpu_verify.py -m gpu -x -q ) > $O/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log; tail -4 $O/pytest_gpu_$TAG.log
timeout 300 python tools/latency_probe.py > $O/latency_$TAG.jsonl 2>$O/latency_$TAG.err; cat $O/latency_$TAG.jsonl; tail -3 $O/latency_$TAG.err
timeout 900 python bench.py --no-cpu-baseline > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"; cat $O/bench_$TAG.json; tail -5 $O/bench_$TAG.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-job-records --slots 1 > $O/bench_under_ncu_$TAG.log 2>&1; echo "launch list exit $?"
python tools/launch_summary.py $O/launches_$TAG.csv > $O/launch_summary_$TAG.txt 2>&1; head -30 $O/launch_summary_$TAG.txt

#!/bin/bash
# round 2, call F: new GPU tests (PoVW, pipelined agent), single-proof latencies with / without the warp-form narrow tree layers,
# launch list of one 2^18 lift
TAG=${1:-r02f}
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log; tail -6 $O/pytest_gpu_$TAG.log
timeout 300 python tools/latency_probe.py > $O/latency_$TAG.jsonl 2>$O/latency_$TAG.err
B200_FOLD_WARP_MAX=0 timeout 300 python tools/latency_probe.py >> $O/latency_$TAG.jsonl 2>>$O/latency_$TAG.err
B200_FOLD_WARP_MAX=1024 timeout 300 python tools/latency_probe.py >> $O/latency_$TAG.jsonl 2>>$O/latency_$TAG.err
B200_FOLD_WARP_MAX=65536 timeout 300 python tools/latency_probe.py >> $O/latency_$TAG.jsonl 2>>$O/latency_$TAG.err
cat $O/latency_$TAG.jsonl; tail -3 $O/latency_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_lift_$TAG.csv python tools/prof_lift.py > $O/prof_lift_$TAG.log 2>&1; echo "lift launch list exit $?"
python tools/launch_summary.py $O/launches_lift_$TAG.csv > $O/launch_summary_lift_$TAG.txt 2>&1; head -40 $O/launch_summary_lift_$TAG.txt

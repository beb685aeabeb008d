#!/bin/bash
# round 2 final (last tree): the GPU suite, smoke and the default bench line
TAG=${1:-r02final3}
O=gpurun_out; mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log; tail -6 $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 $O/smoke_$TAG.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"; cat $O/bench_$TAG.json | cut -c1-300; tail -3 $O/bench_$TAG.err

#!/bin/bash
# round 2, call A: Poseidon2 variant timings, the GPU parity suite with the new default arithmetic, one bench line
TAG=${1:-r02a}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.txt 2>&1
nproc >> $O/smi_$TAG.txt
for f in build/var/mb_*; do echo "== $(basename $f)"; timeout 120 $f | grep -E "kat|fold|perm"; done > $O/mb_variants_$TAG.txt 2>&1
cat $O/mb_variants_$TAG.txt | grep -E "==|1024|FAIL"
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log
tail -4 $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke exit $?"
timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"; cat $O/bench_$TAG.json

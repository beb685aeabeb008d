#!/bin/bash
# round 2, call B: ZALL (all adds on the ALU pipe) variants, launch shapes of the real leaf kernel, one bench line
TAG=${1:-r02b}
O=gpurun_out; mkdir -p $O
for f in build/var/mb_*; do echo "== $(basename $f)"; timeout 120 $f | grep -E "kat|fold|perm"; done > $O/mb_variants_$TAG.txt 2>&1
grep -E "==|1024|FAIL" $O/mb_variants_$TAG.txt
timeout 300 python tools/time_p2.py > $O/time_p2_cfg_$TAG.txt 2>&1; cat $O/time_p2_cfg_$TAG.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_prover.py -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; tail -3 $O/pytest_gpu_$TAG.log
timeout 900 python bench.py --no-cpu-baseline > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"; cat $O/bench_$TAG.json

#!/bin/bash
# round 2, call J: three-pass transform sizes and po2 22-24 segments
TAG=${1:-r02j}
O=gpurun_out; mkdir -p $O
free -g | head -2 > $O/host_mem_$TAG.txt; nvidia-smi --query-gpu=memory.total,memory.used --format=csv >> $O/host_mem_$TAG.txt; cat $O/host_mem_$TAG.txt
( time timeout 1500 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "large_against or three_pass or roundtrip or out_of_range" ) > $O/pytest_ntt_$TAG.log 2>&1
echo "pytest ntt exit $?" >> $O/pytest_ntt_$TAG.log; tail -6 $O/pytest_ntt_$TAG.log
( time timeout 2400 python -m pytest tests/test_gpu_prover.py -m gpu -x -q -k "po2_2" ) > $O/pytest_po2_$TAG.log 2>&1
echo "pytest po2 exit $?" >> $O/pytest_po2_$TAG.log; tail -12 $O/pytest_po2_$TAG.log

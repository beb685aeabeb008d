#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "ntt or expand or intt" > $O/pytest_ntt.log 2>&1; echo "ntt tests exit $?"; tail -5 $O/pytest_ntt.log
timeout 200 python tools/time_ntt2.py > $O/time_ntt2.log 2>&1; cat $O/time_ntt2.log

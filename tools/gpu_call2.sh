#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_compat.py -x -q > $O/pytest_compat.log 2>&1; echo "compat exit $?"; tail -5 $O/pytest_compat.log
timeout 200 python tools/time_p2.py > $O/time_p2.log 2>&1; cat $O/time_p2.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_ntt' -s 5 -c 5 -f -o $O/ncu_ntt_v5 python tools/prof_kernels.py ntt > $O/ncu_ntt_v5.log 2>&1; echo "ncu ntt exit $?"

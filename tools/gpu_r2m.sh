#!/bin/bash
# round 2, call M: full GPU suite, K3 DRAM traffic with / without column batching (ncu --cache-control none: L2 residency is the point)
TAG=${1:-r02m}
O=gpurun_out; mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log; tail -6 $O/pytest_gpu_$TAG.log
B200_NTT_COLBATCH=0 timeout 300 ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_ntt_fwd1|k_ntt_strided_r32' \
    --launch-skip 3 -c 3 --csv --log-file $O/ncu_k3_traffic_b0_$TAG.csv python tools/prof_kernels.py ntt > /dev/null 2>&1; echo "ncu traffic b=0 exit $?"
B200_NTT_COLBATCH=4 timeout 300 ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_ntt_fwd1|k_ntt_strided_r32' \
    --launch-skip 9 -c 9 --csv --log-file $O/ncu_k3_traffic_b4_$TAG.csv python tools/prof_kernels.py ntt > /dev/null 2>&1; echo "ncu traffic b=4 exit $?"
B200_NTT_COLBATCH=2 timeout 300 ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_ntt_fwd1|k_ntt_strided_r32' \
    --launch-skip 17 -c 17 --csv --log-file $O/ncu_k3_traffic_b2_$TAG.csv python tools/prof_kernels.py ntt > /dev/null 2>&1; echo "ncu traffic b=2 exit $?"

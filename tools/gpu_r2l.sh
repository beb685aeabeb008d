#!/bin/bash
# round 2, call L: full GPU suite with the REDUX warp-form permutation, latency probe, bench, K3 DRAM traffic with / without column batching
TAG=${1:-r02l}
O=gpurun_out; mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log; tail -6 $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 $O/smoke_$TAG.log
timeout 300 python tools/latency_probe.py > $O/latency_$TAG.jsonl 2>$O/latency_$TAG.err; cat $O/latency_$TAG.jsonl; tail -3 $O/latency_$TAG.err
timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"; cat $O/bench_$TAG.json | cut -c1-400; tail -5 $O/bench_$TAG.err
B200_NTT_COLBATCH=0 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_ntt_fwd1|k_ntt_strided_r32' \
    --launch-skip 3 -c 3 --csv --log-file $O/ncu_k3_traffic_b0_$TAG.csv python tools/prof_kernels.py ntt > /dev/null 2>&1; echo "ncu traffic b=0 exit $?"
B200_NTT_COLBATCH=4 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_ntt_fwd1|k_ntt_strided_r32' \
    --launch-skip 9 -c 9 --csv --log-file $O/ncu_k3_traffic_b4_$TAG.csv python tools/prof_kernels.py ntt > /dev/null 2>&1; echo "ncu traffic b=4 exit $?"
grep -c k_ntt $O/ncu_k3_traffic_b0_$TAG.csv $O/ncu_k3_traffic_b4_$TAG.csv

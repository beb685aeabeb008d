#!/bin/bash
# round 2, call K: tree with 1 / 2 / 4 Prove tasks in flight on one GPU (4 and 8 segments); K3 DRAM traffic with / without column batching
TAG=${1:-r02k}
O=gpurun_out; mkdir -p $O
for spg in 4 8; do for k in 1 2 3 4; do
  timeout 300 python bench.py --mode tree --segments-per-gpu $spg --tree-seg-in-flight $k 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); c=d['config']; print('spg', c['segments_per_gpu'], 'in_flight', c['prove_tasks_in_flight_per_gpu'], 'ms_to_root %.1f' % c['ms_to_root'], 'seg/s %.2f' % c['segments_per_sec_to_root'])"
done; done | tee $O/tree_inflight_$TAG.txt
timeout 200 python -m pytest tests/test_gpu_prover.py -m gpu -x -q -k "async_job_runner" 2>&1 | tail -2
for b in 0 4; do
B200_NTT_COLBATCH=$b timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_ntt_fwd1|k_ntt_strided_r32<false' \
    --launch-skip $((b == 0 ? 2 : 8)) -c $((b == 0 ? 2 : 8)) --csv --log-file $O/ncu_k3_traffic_b${b}_$TAG.csv python tools/prof_kernels.py ntt > /dev/null 2>&1; echo "ncu traffic b=$b exit $?"
done
head -3 $O/ncu_k3_traffic_b0_$TAG.csv | cut -c1-200

#!/bin/bash
# round 2v: radix-32 strided pass with the tile loaded straight into registers (B200_NTT_R32_DIRECT), A/B + parity + bench
O=gpurun_out; mkdir -p $O
timeout 300 python tools/time_ntt2.py B200_NTT_R32_DIRECT=0 B200_NTT_R32_DIRECT=1 B200_NTT_R32_DIRECT=0 B200_NTT_R32_DIRECT=1 > $O/ntt_r32_direct.txt 2>&1; cat $O/ntt_r32_direct.txt
( time timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_halops.py tests/test_gpu_compat.py -m gpu -x -q ) > $O/pytest_r2v.log 2>&1; grep -E "passed|failed" $O/pytest_r2v.log
for v in 0 1; do
  B200_NTT_R32_DIRECT=$v timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-job-records > $O/bench_r2v_d$v.json 2> $O/bench_r2v_d$v.err; echo "bench direct=$v exit $?"; cut -c1-200 $O/bench_r2v_d$v.json
done

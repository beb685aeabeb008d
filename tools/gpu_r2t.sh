#!/bin/bash
# round 2t: inter-pass twiddle moved from pass A's epilogue to pass B's loads (B200_NTT_TW_IN_B), A/B + parity + bench
O=gpurun_out; mkdir -p $O
timeout 300 python tools/time_ntt2.py B200_NTT_TW_IN_B=0 B200_NTT_TW_IN_B=1 B200_NTT_TW_IN_B=0 B200_NTT_TW_IN_B=1 > $O/ntt_tw_in_b.txt 2>&1; cat $O/ntt_tw_in_b.txt
( time timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_halops.py tests/test_gpu_compat.py -m gpu -x -q ) > $O/pytest_r2t.log 2>&1; tail -4 $O/pytest_r2t.log
for v in 0 1; do
  B200_NTT_TW_IN_B=$v timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-job-records > $O/bench_r2t_tw$v.json 2> $O/bench_r2t_tw$v.err; echo "bench tw_in_b=$v exit $?"; cut -c1-220 $O/bench_r2t_tw$v.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ntt -c 60 --csv --log-file $O/ntt_tw_in_b_launches.csv python tools/time_ntt2.py B200_NTT_TW_IN_B=0 B200_NTT_TW_IN_B=1 > /dev/null 2>&1
python tools/launch_summary.py $O/ntt_tw_in_b_launches.csv 2>/dev/null | head -20

#!/bin/bash
# round 2u: pass A of the inverse transform without the twiddle epilogue: 2 vs 3 CTAs per SM
O=gpurun_out; mkdir -p $O
timeout 300 python tools/time_ntt2.py B200_NTT_TW_IN_B=0 B200_NTT_TW_IN_B=1,B200_NTT_R32_MINB=2 B200_NTT_TW_IN_B=1,B200_NTT_R32_MINB=3 B200_NTT_TW_IN_B=1,B200_NTT_R32_MINB=2 B200_NTT_TW_IN_B=1,B200_NTT_R32_MINB=3 > $O/ntt_tw_in_b2.txt 2>&1; cat $O/ntt_tw_in_b2.txt
( time timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_halops.py tests/test_gpu_compat.py -m gpu -x -q ) > $O/pytest_r2u.log 2>&1; grep -E "passed|failed" $O/pytest_r2u.log

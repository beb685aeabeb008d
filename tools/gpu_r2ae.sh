#!/bin/bash
# round 2ae: per-element tables also in the generic inverse pass B (k_ntt_invb<LOGLC>: check group 2^22, 2^18 proofs); every alternative
# NTT route against the oracle; latencies and headline
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_halops.py tests/test_gpu_compat.py -m gpu -x -q ) > $O/pytest_r2ae.log 2>&1; grep -E "passed|failed" $O/pytest_r2ae.log
timeout 300 python tools/time_ntt2.py B200_NTT_FULL=0 B200_NTT_FULL=1 > $O/ntt_full2.txt 2>&1; cat $O/ntt_full2.txt
timeout 300 python tools/latency_probe.py 2>/dev/null | tee $O/latency_r2ae.jsonl
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-job-records --no-cpu-baseline > $O/b.json 2> $O/b.err; cut -c1-200 $O/b.json

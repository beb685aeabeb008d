#!/bin/bash
# round 2y: warp-per-row radix-32 inverse pass B (k_ntt_invb_r32, B200_NTT_INVB_R32), A/B + parity + bench
O=gpurun_out; mkdir -p $O
timeout 300 python tools/time_ntt2.py B200_NTT_INVB_R32=0 B200_NTT_INVB_R32=1 B200_NTT_INVB_R32=1,B200_NTT_INVB_R32_WAVES=1 B200_NTT_INVB_R32=1,B200_NTT_INVB_R32_WAVES=2 B200_NTT_INVB_R32=0,B200_NTT_INVB_R32_WAVES=0 B200_NTT_INVB_R32=1 > $O/ntt_invb_r32.txt 2>&1; cat $O/ntt_invb_r32.txt
( time timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_halops.py tests/test_gpu_compat.py -m gpu -x -q ) > $O/pytest_r2y.log 2>&1; grep -E "passed|failed" $O/pytest_r2y.log
for v in 0 1; do
  B200_NTT_INVB_R32=$v timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-job-records > $O/b.json 2> $O/b.err
  python - $v >> $O/ntt_invb_r32.txt <<'PY'
import json,sys
d=json.load(open('gpurun_out/b.json')); k=d.get('kernels',[])
print("B200_NTT_INVB_R32=%s value %.3f seg/s  ms/step %.3f  e2e %.3f  K3 %.4f K1 %.4f" % (sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], k[0]['ms'], k[1]['ms']))
PY
done; tail -2 $O/ntt_invb_r32.txt

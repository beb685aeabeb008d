#!/bin/bash
# round 2ab: CTAs per SM of the two radix-32 families (B200_NTT_R32D_MINB 4|5, B200_NTT_INVB_R32_MINB 3|4): standalone, parity, bench
O=gpurun_out; mkdir -p $O
timeout 300 python tools/time_ntt2.py B200_NTT_R32D_MINB=4,B200_NTT_INVB_R32_MINB=3 B200_NTT_R32D_MINB=5,B200_NTT_INVB_R32_MINB=3 B200_NTT_R32D_MINB=4,B200_NTT_INVB_R32_MINB=4 B200_NTT_R32D_MINB=5,B200_NTT_INVB_R32_MINB=4 B200_NTT_R32D_MINB=4,B200_NTT_INVB_R32_MINB=3 B200_NTT_R32D_MINB=5,B200_NTT_INVB_R32_MINB=4 > $O/ntt_minb.txt 2>&1; cat $O/ntt_minb.txt
( time timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q ) > $O/pytest_r2ab.log 2>&1; grep -E "passed|failed" $O/pytest_r2ab.log
for rep in 1 2; do for v in "4 3" "5 4"; do set -- $v
  B200_NTT_R32D_MINB=$1 B200_NTT_INVB_R32_MINB=$2 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-job-records > $O/b.json 2> $O/b.err
  python - "$v" >> $O/ntt_minb.txt <<'PY'
import json,sys
d=json.load(open('gpurun_out/b.json')); k=d.get('kernels',[])
print("R32D_MINB,INVB_R32_MINB=%s value %.3f seg/s  ms/step %.3f  e2e %.3f  K3 %.4f K1 %.4f" % (sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], k[0]['ms'], k[1]['ms']))
PY
done; done; tail -4 $O/ntt_minb.txt

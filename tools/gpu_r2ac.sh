#!/bin/bash
# round 2ac: k_ntt_fwd1<.,2> capped at 40 registers (6 CTAs per SM, B200_NTT_FWD1_MINB=6) against its natural 48 (5 CTAs)
O=gpurun_out; mkdir -p $O
timeout 300 python tools/time_ntt2.py B200_NTT_FWD1_MINB=1 B200_NTT_FWD1_MINB=6 B200_NTT_FWD1_MINB=1 B200_NTT_FWD1_MINB=6 > $O/ntt_fwd1_minb.txt 2>&1
for v in 1 6 1 6; do
  B200_NTT_FWD1_MINB=$v timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-job-records > $O/b.json 2> $O/b.err
  python - "$v" >> $O/ntt_fwd1_minb.txt <<'PY'
import json,sys
d=json.load(open('gpurun_out/b.json')); k=d.get('kernels',[])
print("B200_NTT_FWD1_MINB=%s value %.3f seg/s  ms/step %.3f  e2e %.3f  K3 %.4f K1 %.4f" % (sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], k[0]['ms'], k[1]['ms']))
PY
done; cat $O/ntt_fwd1_minb.txt

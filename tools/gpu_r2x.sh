#!/bin/bash
# round 2x: butterfly adds on the ALU pipe (B200_NTT_ADD_V, compile-time; alternative libraries under build/nttvar/), kernel times,
# parity and the headline bench per variant.  The base library is restored at the end.
O=gpurun_out; mkdir -p $O; : > $O/ntt_addv.txt
cp boundless_b200/libb200zkp.so /tmp/base.so
for v in 0 1 2 3; do
  if [ $v = 0 ]; then cp /tmp/base.so boundless_b200/libb200zkp.so; else cp build/nttvar/libb200zkp_addv$v.so boundless_b200/libb200zkp.so; fi
  echo "== B200_NTT_ADD_V=$v" >> $O/ntt_addv.txt
  timeout 300 python tools/time_ntt2.py B200_NTT_R32_DIRECT=3 B200_NTT_R32_DIRECT=3 >> $O/ntt_addv.txt 2>&1
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "ntt or expand or NTT" 2>&1 | tail -1 >> $O/ntt_addv.txt
  for rep in 1 2; do
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-job-records > $O/b.json 2> $O/b.err
  python - >> $O/ntt_addv.txt <<'PY'
import json
d=json.load(open('gpurun_out/b.json')); k=d.get('kernels',[])
print("value %.3f seg/s  ms/step %.3f  e2e %.3f  K3 %.4f K1 %.4f" % (d['value'], d['ms_per_step'], d['e2e']['value'], k[0]['ms'], k[1]['ms']))
PY
  done
done
cp /tmp/base.so boundless_b200/libb200zkp.so
cat $O/ntt_addv.txt

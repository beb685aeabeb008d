#!/bin/bash
# round 2w: which direction of the radix-32 pass should load straight into registers?  headline bench per setting, two repetitions
O=gpurun_out; mkdir -p $O; : > $O/ntt_r32_direct_bench.txt
for rep in 1 2; do for v in 0 1 2 3; do
  B200_NTT_R32_DIRECT=$v timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-job-records > $O/b.json 2> $O/b.err
  python - $v >> $O/ntt_r32_direct_bench.txt <<'PY'
import json,sys
d=json.load(open('gpurun_out/b.json')); k=d.get('kernels',{})
print("B200_NTT_R32_DIRECT=%s value %.3f seg/s  ms/step %.3f  e2e %.3f  kernels %s" % (sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], json.dumps(k)[:300]))
PY
done; done
cat $O/ntt_r32_direct_bench.txt

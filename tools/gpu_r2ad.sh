#!/bin/bash
# round 2ad: latency form of the warp-cooperative Poseidon2 (B200_P2W_LAT=1, the tree's default) against the first form (alternative
# library build/p2w/libb200zkp_p2wlat0.so): narrow Merkle layers, single-proof latencies, parity, headline
O=gpurun_out; mkdir -p $O; : > $O/p2w_lat.txt
cp boundless_b200/libb200zkp.so /tmp/new.so
for v in old new; do
  if [ $v = old ]; then cp build/p2w/libb200zkp_p2wlat0.so boundless_b200/libb200zkp.so; else cp /tmp/new.so boundless_b200/libb200zkp.so; fi
  echo "== P2Warp $v form" >> $O/p2w_lat.txt
  timeout 120 python tools/time_p2w.py >> $O/p2w_lat.txt 2>&1
  timeout 300 python tools/latency_probe.py 2>/dev/null >> $O/p2w_lat.txt
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-job-records > $O/b.json 2> $O/b.err
  python - >> $O/p2w_lat.txt <<'PY'
import json
d=json.load(open('gpurun_out/b.json'))
print("bench value %.3f seg/s  ms/step %.3f  e2e %.3f" % (d['value'], d['ms_per_step'], d['e2e']['value']))
PY
done
cp /tmp/new.so boundless_b200/libb200zkp.so
cat $O/p2w_lat.txt
( time timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_prover.py tests/test_gpu_verify.py -m gpu -x -q ) > $O/pytest_r2ad.log 2>&1; grep -E "passed|failed" $O/pytest_r2ad.log

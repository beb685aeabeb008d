#!/bin/bash
# 4-slot headline vs the launch shape of the leaf-hash kernel: does leaving register room for other slots' kernels help co-scheduling?
TAG=${1:-r02o}
O=gpurun_out; mkdir -p $O
run() { env "$@" timeout 600 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-job-records 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['kernels']; print('value %.3f seg/s  ms/step %.3f  K4 %.3f ms e2e %.3f' % (d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['e2e']['value']))"; }
for rep in 1 2; do for cfg in 1 0 2 3 4; do echo -n "B200_P2_CFG=$cfg: "; run B200_P2_CFG=$cfg; done; done | tee $O/p2cfg_4slot_$TAG.txt
for sl in 2 3 6; do echo -n "slots=$sl: "; timeout 600 python bench.py --steps 18 --warmup 6 --no-cpu-baseline --no-job-records --slots $sl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value %.3f seg/s  ms/step %.3f' % (d['value'], d['ms_per_step']))"; done | tee -a $O/p2cfg_4slot_$TAG.txt

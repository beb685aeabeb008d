#!/bin/bash
# does the 4-slot headline react to NTT speed?  Same bench with the NTT made slower on purpose (radix-16 strided passes; column batches of 2)
TAG=${1:-r02n}
O=gpurun_out; mkdir -p $O
run() { env "$@" timeout 600 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-job-records 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['kernels']; print('value %.3f seg/s  ms/step %.3f  K3 %.4f ms  K1 %.4f ms  K4 %.3f ms' % (d['value'], d['ms_per_step'], k[0]['ms'], k[1]['ms'], d['roofline']['ms_per_launch']))"; }
for rep in 1 2; do
echo "default:"; run B200_NOP=1
echo "B200_NTT_R32=0:"; run B200_NTT_R32=0
echo "B200_NTT_COLBATCH=2:"; run B200_NTT_COLBATCH=2
echo "slots=1 default:"; env B200_NOP=1 timeout 600 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-job-records --slots 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value %.3f seg/s  ms/step %.3f' % (d['value'], d['ms_per_step']))"
echo "slots=1 B200_NTT_R32=0:"; env B200_NTT_R32=0 timeout 600 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-job-records --slots 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value %.3f seg/s  ms/step %.3f' % (d['value'], d['ms_per_step']))"
done | tee $O/ntt_sensitivity_$TAG.txt

#!/bin/bash
# round 2aa: ncu --set full of the four NTT kernels of the final arrangement (second iteration: tables built, caches as in steady state)
O=gpurun_out; mkdir -p $O
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_ntt' --launch-skip 4 -c 4 -f -o $O/ncu_ntt_r02b \
    python tools/prof_kernels.py ntt2 > $O/ncu_ntt_r02b.log 2>&1; echo "ncu exit $?"
python tools/ncu_brief.py $O/ncu_ntt_r02b.ncu-rep > $O/ncu_brief_r02b_ntt.txt; cat $O/ncu_brief_r02b_ntt.txt

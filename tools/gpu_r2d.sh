#!/bin/bash
# round 2, call D: full GPU suite, the bench line with the queue / tree records, ncu --set full of the NTT passes, the new K4 and the
# latency-form Poseidon2 kernels
TAG=${1:-r02d}
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log; tail -6 $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke exit $?"
timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"; cat $O/bench_$TAG.json; tail -5 $O/bench_$TAG.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_ntt|k_zk' -c 10 -f -o $O/ncu_ntt_$TAG \
    python tools/prof_kernels.py ntt > $O/ncu_ntt_$TAG.log 2>&1; echo "ncu ntt exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_p2_rows' --launch-skip 1 -c 1 -f -o $O/ncu_rows208_$TAG \
    python tools/prof_kernels.py rows208 > $O/ncu_rows208_$TAG.log 2>&1; echo "ncu rows exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_p2_fold_top|k_iop|k_hash' -c 12 -f -o $O/ncu_lat_$TAG \
    python tools/prof_kernels.py tree > $O/ncu_lat_$TAG.log 2>&1; echo "ncu latency kernels exit $?"
ls -la $O/*.ncu-rep

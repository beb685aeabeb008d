#!/bin/bash
# round 2, call S: reference arm at the real size, a long bench line, NVTX-filtered ncu listing, ncu --set full of the round-2 latency kernels
TAG=${1:-r02s}
O=gpurun_out; mkdir -p $O
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_reference_$TAG.json 2> $O/bench_reference_$TAG.err; echo "reference exit $?"; cat $O/bench_reference_$TAG.json | cut -c1-900; tail -4 $O/bench_reference_$TAG.err
timeout 900 python bench.py --steps 48 --warmup 8 > $O/bench_long_$TAG.json 2> $O/bench_long_$TAG.err; echo "bench exit $?"; cat $O/bench_long_$TAG.json | cut -c1-600
timeout 300 ncu --nvtx --nvtx-include "b200/verify_integrity/" --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/ncu_nvtx_verify_$TAG.csv \
    python tools/prof_lift.py > $O/ncu_nvtx_$TAG.log 2>&1; echo "ncu nvtx exit $?"; grep -c "gpu__time_duration" $O/ncu_nvtx_verify_$TAG.csv
timeout 300 ncu --nvtx --nvtx-include "FRI commit (K3, K4, K5, K6)/" --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/ncu_nvtx_fri_$TAG.csv \
    python tools/prof_lift.py > $O/ncu_nvtx2_$TAG.log 2>&1; echo "ncu nvtx fri exit $?"; grep -c "gpu__time_duration" $O/ncu_nvtx_fri_$TAG.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_p2_fold_w|k_ntt_row_twiddle|k_verify_queries|k_iop_commit_elems' -c 8 -f -o $O/ncu_r2kernels_$TAG \
    python tools/prof_kernels.py tree > $O/ncu_r2kernels_$TAG.log 2>&1; echo "ncu r2 kernels exit $?"

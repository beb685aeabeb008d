#!/bin/bash
# round 2, call I: new GPU tests, compute-sanitizer on the round-2 code paths, column batching of expand+NTT (time + DRAM traffic)
TAG=${1:-r02i}
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_tasks.py -m gpu -x -q ) > $O/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log; tail -4 $O/pytest_gpu_$TAG.log
SEL="prove_lift_composite or join_over_device or verify_integrity_binds or async_job_runner or povw_kinds or keccak"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_prover.py -m gpu -x -q -k "$SEL" > $O/sanitizer_memcheck_$TAG.txt 2>&1
echo "memcheck exit $?"; tail -5 $O/sanitizer_memcheck_$TAG.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_prover.py tests/test_gpu_kernels.py -m gpu -x -q -k "prove_lift_composite or join_over_device or merkle_tree and not full" > $O/sanitizer_racecheck_$TAG.txt 2>&1
echo "racecheck exit $?"; tail -5 $O/sanitizer_racecheck_$TAG.txt
for b in 0 8 4 2; do B200_NTT_COLBATCH=$b timeout 120 python tools/time_ntt2.py "B200_NTT_R32=1" 2>&1 | sed "s/^/colbatch $b: /"; done > $O/ntt_colbatch_$TAG.txt; cat $O/ntt_colbatch_$TAG.txt
for b in 0 4; do
B200_NTT_COLBATCH=$b timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_ntt_fwd1|k_ntt_strided_r32' \
    --launch-skip 10 -c 10 --csv --log-file $O/ncu_k3_traffic_b${b}_$TAG.csv python tools/prof_kernels.py ntt > /dev/null 2>&1; echo "ncu traffic b=$b exit $?"
done
python - <<'PY'
import csv, sys
for b in (0, 4):
    rows = [r for r in csv.reader(open("gpurun_out/ncu_k3_traffic_b%d_%s.csv" % (b, sys.argv[1] if len(sys.argv) > 1 else "r02i"))) if len(r) > 5]
    hdr = rows[0]; ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    tot = {}
    for r in rows[1:]:
        name = r[ki].split("(")[0][:40]
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        if "byte" in u.lower():
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            tot[name + " bytes"] = tot.get(name + " bytes", 0) + v
        else:
            tot[name + " time " + u] = tot.get(name + " time " + u, 0) + v
    print("colbatch", b, {k: round(v / 1e6, 2) if "bytes" in k else round(v, 1) for k, v in tot.items()})
PY

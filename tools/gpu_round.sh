#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the bench line, the ncu launch list of the bench command and ncu --set full
# captures of the standalone K5-K8 kernels.  Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/ by hand.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu_$TAG.log
tail -3 $O/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke exit $?"
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"; cat $O/bench_$TAG.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --slots 1 > $O/bench_under_ncu_$TAG.log 2>&1; echo "launch list exit $?"
python tools/launch_summary.py $O/launches_$TAG.csv > $O/launch_summary_$TAG.txt 2>&1; head -12 $O/launch_summary_$TAG.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_fri_fold|k_ev_|k_mix_poly|k_eltwise_sum|k_deep' \
    -c 12 -f -o $O/ncu_hal_$TAG python tools/prof_kernels.py hal > $O/ncu_hal_$TAG.log 2>&1; echo "ncu hal exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_p2_fold' -c 1 -f -o $O/ncu_fold_$TAG \
    python tools/prof_kernels.py fold > $O/ncu_fold_$TAG.log 2>&1; echo "ncu fold exit $?"

timeout 300 python tools/ntt_sweep.py > $O/ntt_sweep_$TAG.jsonl 2> $O/ntt_sweep_$TAG.err; echo "sweep exit $?"
timeout 120 python tools/time_ntt2.py > $O/time_ntt2_$TAG.log 2>&1

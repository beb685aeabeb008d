#!/bin/bash
O=gpurun_out; mkdir -p $O
run(){ python bench.py --no-cpu-baseline "$@" 2>>$O/bench_ab.err | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$TAGX', d['value'],d['e2e']['value'],d['ms_per_step'])"; }
for r in 0 1 0 1; do TAGX="R32=$r slots=4"; B200_NTT_R32=$r run --slots 4; done
for r in 0 1; do TAGX="R32=$r slots=1"; B200_NTT_R32=$r run --slots 1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_v8.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --slots 1 > $O/bench_under_ncu_v8.log 2>&1
python tools/launch_summary.py $O/launches_v8.csv > $O/launch_summary_v8.txt 2>&1; head -14 $O/launch_summary_v8.txt

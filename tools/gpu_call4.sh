#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_v8.log 2>&1; echo "gpu tests exit $?"; tail -3 $O/pytest_gpu_v8.log
timeout 300 python bench.py --no-cpu-baseline > $O/bench_v8_s4.json 2>$O/bench_v8.err; echo "bench exit $?"; python -c "import json;d=json.load(open('$O/bench_v8_s4.json'));print(d['value'],d['e2e']['value'],d['ms_per_step'])"
timeout 300 python bench.py --no-cpu-baseline --slots 8 > $O/bench_v8_s8.json 2>>$O/bench_v8.err; echo "bench exit $?"; python -c "import json;d=json.load(open('$O/bench_v8_s8.json'));print(d['value'],d['e2e']['value'],d['ms_per_step'])"
timeout 300 python bench.py --no-cpu-baseline --slots 2 > $O/bench_v8_s2.json 2>>$O/bench_v8.err; echo "bench exit $?"; python -c "import json;d=json.load(open('$O/bench_v8_s2.json'));print(d['value'],d['e2e']['value'],d['ms_per_step'])"

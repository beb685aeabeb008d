#!/bin/bash
# N ranks under torchrun exactly as the driver launches them: reference arm skipped (CPU), our arm at N, plus tree mode with 4 segments/GPU
TAG=${1:-r02}
N=${2:-8}
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/topo_${N}gpu_$TAG.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 \
    > $O/bench_${N}gpu_$TAG.json 2> $O/bench_${N}gpu_$TAG.err; echo "bench exit $?"; cat $O/bench_${N}gpu_$TAG.json; tail -5 $O/bench_${N}gpu_$TAG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --mode tree --segments-per-gpu 8 \
    > $O/bench_tree_${N}gpu_spg8_$TAG.json 2> $O/bench_tree_${N}gpu_spg8_$TAG.err; echo "tree exit $?"; cat $O/bench_tree_${N}gpu_spg8_$TAG.json; tail -3 $O/bench_tree_${N}gpu_spg8_$TAG.err

#!/bin/bash
# two ranks under torchrun: the contract line (with queue / tree records) and the tree mode alone (config 4)
TAG=${1:-r02}
N=${2:-2}
O=gpurun_out; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 12 --warmup 3 \
    > $O/bench_${N}gpu_$TAG.json 2> $O/bench_${N}gpu_$TAG.err; echo "bench exit $?"; cat $O/bench_${N}gpu_$TAG.json; tail -5 $O/bench_${N}gpu_$TAG.err
for spg in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --mode tree --segments-per-gpu $spg \
    > $O/bench_tree_${N}gpu_spg${spg}_$TAG.json 2> $O/bench_tree_${N}gpu_spg${spg}_$TAG.err; echo "tree exit $?"; cat $O/bench_tree_${N}gpu_spg${spg}_$TAG.json; tail -3 $O/bench_tree_${N}gpu_spg${spg}_$TAG.err
done

#!/bin/bash
# 2-GPU check of the contract launch (torchrun, one rank per GPU over NCCL): segments mode and the join-tree job (BASELINE config 4)
O=gpurun_out; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench 2gpu exit $?"; tail -1 $O/bench_2gpu.json | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --mode tree --segments-per-gpu 4 > $O/bench_tree_2gpu.json 2> $O/bench_tree_2gpu.err; echo "tree 2gpu exit $?"; tail -1 $O/bench_tree_2gpu.json | cut -c1-400

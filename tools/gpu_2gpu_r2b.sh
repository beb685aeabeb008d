#!/bin/bash
# two ranks under torchrun on the final round-2 tree: the contract line with its queue / tree records (CPU arm skipped: it is rank 0 alone)
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu-baseline \
    > $O/bench_2gpu_r02b.json 2> $O/bench_2gpu_r02b.err; echo "bench exit $?"; cat $O/bench_2gpu_r02b.json; tail -3 $O/bench_2gpu_r02b.err

#!/bin/bash
# builds build/microbench_p2v{V} for the Poseidon2 linear-layer add variants (bitmask B200_P2_V, see csrc/poseidon2.cuh)
set -e
cd "$(dirname "$0")/.."
mkdir -p build
for V in 0 1 2 4 8 3 5 7 15 12; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -ccbin /usr/bin/g++ -DB200_P2_V=$V \
     -o build/microbench_p2v${V} boundless_b200/tools/microbench.cu &
done; wait

#!/bin/bash
# builds build/microbench_a{A}_r{R} for the field-op variants (see csrc/field.cuh)
set -e
cd "$(dirname "$0")/.."
mkdir -p build
for A in 0 1; do for R in 0 1 2; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -ccbin /usr/bin/g++ -DB200_ADD_V=$A -DB200_REDC_V=$R \
     -o build/microbench_a${A}_r${R} boundless_b200/tools/microbench.cu &
done; done; wait

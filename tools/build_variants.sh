#!/bin/bash
# Builds quick-mode microbenchmarks (KAT + 6006-permutation differential fold + four launch shapes of the Poseidon2 permutation) for
# compile-time variants of csrc/poseidon2.cuh / csrc/field.cuh into build/var/, to be run on the GPU by tools/gpu_mb.sh.
#   tools/build_variants.sh "tag:-Dflags" ...      e.g.  "l7:-DB200_P2_LAZY=7"  "r3:-DB200_REDC_V=3"
set -e
cd "$(dirname "$0")/.."
mkdir -p build/var
NVCC="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -ccbin /usr/bin/g++ -DB200_MB_QUICK"
n=0
for v in "$@"; do
  tag=${v%%:*}; fl=${v#*:}
  $NVCC $fl -o build/var/mb_$tag boundless_b200/tools/microbench.cu &
  n=$((n+1)); if [ $((n % 8)) -eq 0 ]; then wait; fi
done
wait
ls build/var

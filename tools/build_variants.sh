#!/bin/bash
# Builds quick-mode microbenchmarks (KAT + four launch shapes of the Poseidon2 permutation) for the compile-time variants of
# csrc/poseidon2.cuh / csrc/field.cuh into build/var/, to be run on the GPU by tools/gpu_mb.sh:
#   B200_P2_LAZY  1 lazy x^4, 2 64-bit sum of the internal layer, 4 lazy internal state  (prepared, host-validated, not yet timed)
#   B200_P2_Z     three-input adds (measured: neutral),  B200_P2_V VIADDMNMX-only adds (measured: slower),  B200_REDC_V  reductions
set -e
cd "$(dirname "$0")/.."
mkdir -p build/var
NVCC="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -ccbin /usr/bin/g++ -DB200_MB_QUICK"
for L in 0 1 2 3 4 5 6 7; do $NVCC -DB200_P2_LAZY=$L -o build/var/mb_lazy$L boundless_b200/tools/microbench.cu & done; wait
for extra in "$@"; do   # e.g. tools/build_variants.sh "-DB200_P2_Z=16 -DB200_P2_LAZY=7"
  tag=$(echo "$extra" | tr -cd 'A-Za-z0-9=_' | tr '=' '_'); $NVCC $extra -o build/var/mb_$tag boundless_b200/tools/microbench.cu
done
ls build/var

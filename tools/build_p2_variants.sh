#!/bin/bash
# tools/build_p2_variants.sh "tag:-Dflags" ...   -> build/p2var/libp2_<tag>.so  (csrc/hash.cu only; see tools/p2only.cu)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/p2var
NVCC="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC -shared -cudart static -I boundless_b200/csrc -I include"
n=0
for v in "$@"; do
  tag=${v%%:*}; fl=${v#*:}
  $NVCC $fl -o build/p2var/libp2_$tag.so tools/p2only.cu &
  n=$((n+1)); if [ $((n % 8)) -eq 0 ]; then wait; fi
done
wait
ls build/p2var

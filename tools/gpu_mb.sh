#!/bin/bash
# run every microbench variant under build/var (quick mode) and collect the permutation rates
O=gpurun_out; mkdir -p $O
for f in build/var/mb_*; do echo "== $(basename $f)"; timeout 60 $f | grep -E "kat|perm"; done > $O/mb_variants.txt 2>&1
cat $O/mb_variants.txt

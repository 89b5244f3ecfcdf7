#!/bin/bash
mkdir -p gpurun_out
for v in trall trnone trall_mb8 trnone_mb8; do
  BSA_LIB_PATH=$PWD/tools/microbench/libbsa_$v.so BSA_PROFILE_GROUPS=1 python tools/quick_bench.py 10000 0 > gpurun_out/c4_groups_$v.txt 2>&1
done
echo done

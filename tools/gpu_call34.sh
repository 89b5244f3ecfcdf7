#!/bin/bash
mkdir -p gpurun_out
for v in cur unaligned cur unaligned; do
  if [ "$v" = cur ]; then python tools/quick_ovm.py 1000 50000 2>&1 | sed "s/^/[$v] /"; else BSA_LIB_PATH=$PWD/tools/microbench/libbsa_$v.so python tools/quick_ovm.py 1000 50000 2>&1 | sed "s/^/[$v] /"; fi
done > gpurun_out/c34_ab_ovm.txt 2>&1
cat gpurun_out/c34_ab_ovm.txt
BSA_PROFILE_GROUPS=1 python tools/quick_ovm.py 1000 50000 2>&1 | grep "group16" | tail -45 > gpurun_out/c34_groups16.txt
python tools/quick_bench.py 10000 2 | tail -1

#!/bin/bash
mkdir -p gpurun_out
for v in cur noswap cur noswap; do
  if [ "$v" = cur ]; then python tools/quick_ovm.py 1000 50000 2>&1 | head -1 | sed "s/^/[$v] /"; else BSA_NO_SWAP16=1 python tools/quick_ovm.py 1000 50000 2>&1 | head -1 | sed "s/^/[$v] /"; fi
done > gpurun_out/c36_ab_swap16_ovm.txt 2>&1
cat gpurun_out/c36_ab_swap16_ovm.txt
BSA_PROFILE_GROUPS=1 python tools/quick_ovm.py 1000 50000 2>&1 | grep "group16 B" | head -20

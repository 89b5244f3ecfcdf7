#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/c35_pytest.txt
cat gpurun_out/c35_pytest.txt
for v in cur noswap cur noswap; do
  if [ "$v" = cur ]; then python tools/quick_ovm.py 1000 50000 2>&1 | head -1 | sed "s/^/[$v] /"; else BSA_NO_SWAP16=1 python tools/quick_ovm.py 1000 50000 2>&1 | head -1 | sed "s/^/[$v] /"; fi
done > gpurun_out/c35_ab_swap16_ovm.txt 2>&1
cat gpurun_out/c35_ab_swap16_ovm.txt
BSA_PROFILE_GROUPS=1 python tools/quick_ovm.py 1000 50000 2>&1 | grep "group16\|bsa group\|bsa pair" | head -80 > gpurun_out/c35_groups16.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "score_only or alphabet" 2>&1 | tail -3
for v in quad noquad quad noquad; do
  if [ $v = quad ]; then unset BSA_NO_QUAD16; else export BSA_NO_QUAD16=1; fi
  echo "[$v] $(python tools/quick_ovm.py 1000 50000 2>&1 | head -1)"
done > gpurun_out/c62_ab_quad16.txt 2>&1
cat gpurun_out/c62_ab_quad16.txt
unset BSA_NO_QUAD16
BSA_PROFILE_GROUPS=1 python tools/quick_ovm.py 1000 50000 2>&1 | grep "quad16" | head -20

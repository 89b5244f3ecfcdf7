#!/bin/bash
mkdir -p gpurun_out
for g in 148 144 148 144; do
  echo "[grid $g] $(BSA_WAVE_GRID=$g BSA_CFG5_REPS=25 BSA_CFG5_NOCHECK=1 python tools/cfg5_run.py 2>&1 | tail -1 | grep -o '"kernel_ms_all_reps": [^]]*]')"
done > gpurun_out/c21_grid.txt 2>&1
cat gpurun_out/c21_grid.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_local.py -m gpu -x -q 2>&1 | tail -3

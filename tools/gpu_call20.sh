#!/bin/bash
mkdir -p gpurun_out
for g in 148 144 140 148 144 140; do
  echo "[grid $g] $(BSA_WAVE_GRID=$g BSA_CFG5_REPS=25 BSA_CFG5_NOCHECK=1 python tools/cfg5_run.py 2>&1 | tail -1 | grep -o '"kernel_ms_all_reps": [^]]*]')"
done > gpurun_out/c20_grid.txt 2>&1
cat gpurun_out/c20_grid.txt
nvidia-smi --query-compute-apps=pid,used_memory --format=csv

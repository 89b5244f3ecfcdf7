#!/bin/bash
mkdir -p gpurun_out
for round in 1 2; do BSA_CFG5_REPS=8 BSA_CFG5_NOCHECK=1 timeout 300 python tools/cfg5_run.py 2>&1 | tail -1; done > gpurun_out/c17_cfg5.txt
grep -o '"kernel_ms_all_reps": [^]]*]' gpurun_out/c17_cfg5.txt
BSA_WAVE_TRACE=gpurun_out/c17_wave_trace.csv BSA_CFG5_NOCHECK=1 timeout 300 python tools/cfg5_run.py 2>&1 | tail -1 | cut -c60-200
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv

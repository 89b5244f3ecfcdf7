#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu,utilization.gpu --format=csv,noheader
BSA_CFG5_REPS=25 BSA_CFG5_NOCHECK=1 tools/ab_run.sh "python tools/cfg5_run.py" cur nohi afma > gpurun_out/c19_ab.txt 2>&1
grep -o '^\[[a-z]*\]\|"kernel_ms_all_reps": [^]]*]' gpurun_out/c19_ab.txt | paste - -

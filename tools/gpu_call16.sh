#!/bin/bash
mkdir -p gpurun_out
{
python tools/wave_probe.py 34350 8704 1
python tools/wave_probe.py 34350 35000 1
python tools/wave_probe.py 34350 35000 4
} > gpurun_out/c16_probe.txt 2>&1
cat gpurun_out/c16_probe.txt
for round in 1 2 3; do BSA_CFG5_NOCHECK=1 timeout 300 python tools/cfg5_run.py 2>&1 | tail -1; done > gpurun_out/c16_cfg5.txt
cut -c60-200 gpurun_out/c16_cfg5.txt
BSA_CFG5_NOCHECK=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c16_launches_cfg5.csv python tools/cfg5_run.py > /dev/null 2>&1
grep -E "gotoh_wave|traceback" gpurun_out/c16_launches_cfg5.csv | awk -F'","' '{print $5, $NF}'
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_callers_gpu.py tests/test_gpu_zconfigs_at_size.py -m gpu -x -q -k "wave or titin or cfg5" 2>&1 | tail -3

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_callers_gpu.py tests/test_gpu_zconfigs_at_size.py -m gpu -x -q -k "wave or titin or cfg5" 2>&1 | tail -5 > gpurun_out/c14_pytest.txt
cat gpurun_out/c14_pytest.txt
for round in 1 2 3; do BSA_CFG5_NOCHECK=1 timeout 300 python tools/cfg5_run.py 2>&1 | tail -1; done > gpurun_out/c14_cfg5.txt
cut -c60-200 gpurun_out/c14_cfg5.txt
BSA_CFG5_NOCHECK=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c14_launches_cfg5.csv python tools/cfg5_run.py > /dev/null 2>&1
grep -E "gotoh_wave|traceback" gpurun_out/c14_launches_cfg5.csv | awk -F'","' '{print $5, $NF}'
BSA_WAVE_TRACE=gpurun_out/c14_wave_trace.csv BSA_CFG5_NOCHECK=1 timeout 300 python tools/cfg5_run.py 2>&1 | tail -1 | cut -c60-200
echo done

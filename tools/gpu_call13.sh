#!/bin/bash
mkdir -p gpurun_out
BSA_CFG5_NOCHECK=1 tools/ab_run.sh "python tools/cfg5_run.py" cur nohi sl2k sl128 > gpurun_out/c13_ab_wave.txt 2>&1
cat gpurun_out/c13_ab_wave.txt | cut -c1-40,100-200
BSA_CFG5_NOCHECK=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c13_launches_cfg5.csv python tools/cfg5_run.py > /dev/null 2>&1
grep -E "gotoh_wave|traceback" gpurun_out/c13_launches_cfg5.csv | awk -F'","' '{print $5, $NF}'
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c13_pytest.txt
cat gpurun_out/c13_pytest.txt
echo done

#!/bin/bash
# new K3 (direction-frame cell, top padding) + run-parallel traceback: parity, timing, launch list, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_callers_gpu.py tests/test_gpu_zconfigs_at_size.py -m gpu -x -q -k "wave or path or titin or long or cfg5 or traceback or replay or protocol" 2>&1 | tail -15 > gpurun_out/c12_pytest.txt
cat gpurun_out/c12_pytest.txt
for round in 1 2; do BSA_CFG5_NOCHECK=1 timeout 300 python tools/cfg5_run.py 2>&1 | tail -1; done > gpurun_out/c12_cfg5.txt
cat gpurun_out/c12_cfg5.txt
BSA_CFG5_NOCHECK=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c12_launches_cfg5.csv python tools/cfg5_run.py > /dev/null 2>&1
tail -8 gpurun_out/c12_launches_cfg5.csv | cut -c1-40,200-400
export BSA_CFG5_NOCHECK=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gotoh_wave -s 1 -c 1 -o gpurun_out/c12_prof_wave python tools/cfg5_run.py > gpurun_out/c12_ncu.log 2>&1
ncu -i gpurun_out/c12_prof_wave.ncu-rep --page raw --csv > gpurun_out/c12_prof_wave_raw.csv 2>/dev/null
ncu -i gpurun_out/c12_prof_wave.ncu-rep --page source --csv > gpurun_out/c12_prof_wave_source.csv 2>/dev/null
rm -f gpurun_out/c12_prof_wave.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traceback_kernel -s 1 -c 1 -o gpurun_out/c12_prof_tb python tools/cfg5_run.py > gpurun_out/c12_ncu2.log 2>&1
ncu -i gpurun_out/c12_prof_tb.ncu-rep --page raw --csv > gpurun_out/c12_prof_tb_raw.csv 2>/dev/null
ncu -i gpurun_out/c12_prof_tb.ncu-rep --page source --csv > gpurun_out/c12_prof_tb_source.csv 2>/dev/null
rm -f gpurun_out/c12_prof_tb.ncu-rep
echo done

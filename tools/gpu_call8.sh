#!/bin/bash
mkdir -p gpurun_out
export BSA_CFG5_NOCHECK=1
ncu --set full --clock-control none --import-source on -k regex:gotoh_wave -s 1 -c 1 -o gpurun_out/c8_prof_wave2 python tools/cfg5_run.py > gpurun_out/c8_ncu.log 2>&1
ncu -i gpurun_out/c8_prof_wave2.ncu-rep --page raw --csv > gpurun_out/c8_prof_wave2_raw.csv 2>/dev/null
ncu -i gpurun_out/c8_prof_wave2.ncu-rep --page source --csv > gpurun_out/c8_prof_wave2_source.csv 2>/dev/null
rm -f gpurun_out/c8_prof_wave2.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:traceback_kernel -s 1 -c 1 -o gpurun_out/c8_prof_tb python tools/cfg5_run.py > gpurun_out/c8_ncu2.log 2>&1
ncu -i gpurun_out/c8_prof_tb.ncu-rep --page raw --csv > gpurun_out/c8_prof_tb_raw.csv 2>/dev/null
ncu -i gpurun_out/c8_prof_tb.ncu-rep --page source --csv > gpurun_out/c8_prof_tb_source.csv 2>/dev/null
rm -f gpurun_out/c8_prof_tb.ncu-rep
echo done

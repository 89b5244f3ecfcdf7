#!/bin/bash
mkdir -p gpurun_out
export BSA_CFG5_NOCHECK=1
ncu --set full --clock-control none --import-source on -k regex:gotoh_wave -s 1 -c 1 -o gpurun_out/c10_prof_wave python tools/cfg5_run.py > gpurun_out/c10_ncu.log 2>&1
ncu -i gpurun_out/c10_prof_wave.ncu-rep --page raw --csv > gpurun_out/c10_prof_wave_raw.csv 2>/dev/null
ncu -i gpurun_out/c10_prof_wave.ncu-rep --page source --csv > gpurun_out/c10_prof_wave_source.csv 2>/dev/null
rm -f gpurun_out/c10_prof_wave.ncu-rep
echo done

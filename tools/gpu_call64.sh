#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gotoh_score16_quad_kernel -s 24 -c 1 -o gpurun_out/c64_prof_quad python tools/quick_ovm.py 1000 50000 > gpurun_out/c64_ncu_quad.log 2>&1
ncu -i gpurun_out/c64_prof_quad.ncu-rep --page raw --csv > gpurun_out/c64_prof_quad_raw.csv 2>/dev/null
ncu -i gpurun_out/c64_prof_quad.ncu-rep --page source --csv > gpurun_out/c64_prof_quad_source.csv 2>/dev/null
rm -f gpurun_out/c64_prof_quad.ncu-rep
tail -2 gpurun_out/c64_ncu_quad.log

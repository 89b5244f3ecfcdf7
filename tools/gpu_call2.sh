#!/bin/bash
mkdir -p gpurun_out
python tools/flag_cost_probe.py > gpurun_out/c2_flag_cost.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:gotoh_stream_kernel -s 16 -c 1 -o gpurun_out/c2_prof_stream17 python tools/quick_bench.py 10000 0 > gpurun_out/c2_ncu1.log 2>&1
ncu -i gpurun_out/c2_prof_stream17.ncu-rep --page raw --csv > gpurun_out/c2_prof_stream17_raw.csv 2>/dev/null
ncu -i gpurun_out/c2_prof_stream17.ncu-rep --page source --csv > gpurun_out/c2_prof_stream17_source.csv 2>/dev/null
ls -la gpurun_out
echo done

#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gotoh_stream_kernel -s 16 -c 1 -o gpurun_out/c31_prof_stream python tools/quick_bench.py 10000 0 > gpurun_out/c31_ncu1.log 2>&1
ncu -i gpurun_out/c31_prof_stream.ncu-rep --page raw --csv > gpurun_out/c31_prof_stream_raw.csv 2>/dev/null
ncu -i gpurun_out/c31_prof_stream.ncu-rep --page source --csv > gpurun_out/c31_prof_stream_source.csv 2>/dev/null
rm -f gpurun_out/c31_prof_stream.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gotoh_pair_kernel -s 3 -c 1 -o gpurun_out/c31_prof_pair python tools/quick_bench.py 10000 0 > gpurun_out/c31_ncu2.log 2>&1
ncu -i gpurun_out/c31_prof_pair.ncu-rep --page raw --csv > gpurun_out/c31_prof_pair_raw.csv 2>/dev/null
ncu -i gpurun_out/c31_prof_pair.ncu-rep --page source --csv > gpurun_out/c31_prof_pair_source.csv 2>/dev/null
rm -f gpurun_out/c31_prof_pair.ncu-rep
echo done

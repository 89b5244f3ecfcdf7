#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --single-process > gpurun_out/c23_bench_cfg2_sp.json 2> gpurun_out/c23_sp.err
python -c "
import json;d=json.loads(open('gpurun_out/c23_bench_cfg2_sp.json').read().strip().split('\n')[-1]);print('single-process', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c23_launches_cfg2.csv python tools/quick_bench.py 10000 0 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gotoh_stream_kernel -s 16 -c 1 -o gpurun_out/c23_prof_stream python tools/quick_bench.py 10000 0 > gpurun_out/c23_ncu1.log 2>&1
ncu -i gpurun_out/c23_prof_stream.ncu-rep --page raw --csv > gpurun_out/c23_prof_stream_raw.csv 2>/dev/null
ncu -i gpurun_out/c23_prof_stream.ncu-rep --page source --csv > gpurun_out/c23_prof_stream_source.csv 2>/dev/null
rm -f gpurun_out/c23_prof_stream.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gotoh_pair_kernel -s 3 -c 1 -o gpurun_out/c23_prof_pair python tools/quick_bench.py 10000 0 > gpurun_out/c23_ncu2.log 2>&1
ncu -i gpurun_out/c23_prof_pair.ncu-rep --page raw --csv > gpurun_out/c23_prof_pair_raw.csv 2>/dev/null
ncu -i gpurun_out/c23_prof_pair.ncu-rep --page source --csv > gpurun_out/c23_prof_pair_source.csv 2>/dev/null
rm -f gpurun_out/c23_prof_pair.ncu-rep
BSA_PROFILE_GROUPS=1 python tools/quick_bench.py 10000 1 > gpurun_out/c23_groups.txt 2>&1
echo done

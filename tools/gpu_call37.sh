#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c37_bench_cfg2.json 2> gpurun_out/c37_bench_cfg2.err
tail -1 gpurun_out/c37_bench_cfg2.json | cut -c1-400
timeout 900 python bench.py --steps 3 --warmup 3 --workload cfg4 > gpurun_out/c37_bench_cfg4.json 2> gpurun_out/c37_bench_cfg4.err
tail -1 gpurun_out/c37_bench_cfg4.json | cut -c1-400
for k in stream pair; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gotoh_${k}_kernel -s 16 -c 1 -o gpurun_out/c37_prof_$k python tools/quick_bench.py 10000 0 > gpurun_out/c37_ncu_$k.log 2>&1
ncu -i gpurun_out/c37_prof_$k.ncu-rep --page raw --csv > gpurun_out/c37_prof_${k}_raw.csv 2>/dev/null
ncu -i gpurun_out/c37_prof_$k.ncu-rep --page source --csv > gpurun_out/c37_prof_${k}_source.csv 2>/dev/null
rm -f gpurun_out/c37_prof_$k.ncu-rep
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gotoh_score16_kernel -s 14 -c 1 -o gpurun_out/c37_prof_s16 python tools/quick_ovm.py 1000 20000 > gpurun_out/c37_ncu_s16.log 2>&1
ncu -i gpurun_out/c37_prof_s16.ncu-rep --page raw --csv > gpurun_out/c37_prof_s16_raw.csv 2>/dev/null
ncu -i gpurun_out/c37_prof_s16.ncu-rep --page source --csv > gpurun_out/c37_prof_s16_source.csv 2>/dev/null
rm -f gpurun_out/c37_prof_s16.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c37_launches_cfg2.csv python tools/quick_bench.py 10000 0 > /dev/null 2>&1
echo done

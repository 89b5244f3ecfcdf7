#!/bin/bash
mkdir -p gpurun_out
for w in cfg2 cfg1 cfg4 cfg5 cfg3shard; do
  timeout 1200 python bench.py --steps 3 --warmup 3 --workload $w > gpurun_out/c52_bench_$w.json 2> gpurun_out/c52_bench_$w.err
  python -c "
import json;d=json.loads(open('gpurun_out/c52_bench_$w.json').read().strip().split('\n')[-1]);print('$w', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'alu', round(d['roofline']['alu_pipe']['frac'],3), 'peak', round(d['roofline']['peak'],2), round(d['roofline']['alu_pipe']['peak'],2), 'cpu', d.get('cpu_baseline',{}).get('value'), 'launches', d['gpu_launches'])"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c52_bench_reference.json 2> gpurun_out/c52_bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c52_smoke.txt 2>&1; tail -1 gpurun_out/c52_smoke.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gotoh_stream_kernel -s 16 -c 1 -o gpurun_out/c52_prof_stream python tools/quick_bench.py 10000 0 > gpurun_out/c52_ncu_stream.log 2>&1
ncu -i gpurun_out/c52_prof_stream.ncu-rep --page raw --csv > gpurun_out/c52_prof_stream_raw.csv 2>/dev/null
ncu -i gpurun_out/c52_prof_stream.ncu-rep --page source --csv > gpurun_out/c52_prof_stream_source.csv 2>/dev/null
rm -f gpurun_out/c52_prof_stream.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gotoh_score16_kernel -s 12 -c 1 -o gpurun_out/c52_prof_s16 python tools/quick_ovm.py 1000 50000 > gpurun_out/c52_ncu_s16.log 2>&1
ncu -i gpurun_out/c52_prof_s16.ncu-rep --page raw --csv > gpurun_out/c52_prof_s16_raw.csv 2>/dev/null
ncu -i gpurun_out/c52_prof_s16.ncu-rep --page source --csv > gpurun_out/c52_prof_s16_source.csv 2>/dev/null
rm -f gpurun_out/c52_prof_s16.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c52_launches_cfg2.csv python tools/quick_bench.py 10000 0 > /dev/null 2>&1
echo done

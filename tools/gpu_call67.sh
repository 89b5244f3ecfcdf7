#!/bin/bash
mkdir -p gpurun_out
timeout 140 python bench.py > gpurun_out/c67_bench_default.json 2> gpurun_out/c67_bench_default.err
python -c "
import json;d=json.loads(open('gpurun_out/c67_bench_default.json').read().strip().split('\n')[-1]);print('default', round(d['value'],1), 'steps', d['steps'], 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'cpu', d.get('cpu_baseline',{}).get('value'), d['clocks'])"

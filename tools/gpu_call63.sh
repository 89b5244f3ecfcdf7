#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/c63_pytest.txt; cat gpurun_out/c63_pytest.txt
timeout 900 python bench.py --steps 3 --warmup 3 --workload cfg4 > gpurun_out/c63_bench_cfg4.json 2> gpurun_out/c63_bench_cfg4.err
python -c "
import json;d=json.loads(open('gpurun_out/c63_bench_cfg4.json').read().strip().split('\n')[-1]);print('cfg4', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'alu', round(d['roofline']['alu_pipe']['frac'],3), 'cpu', d.get('cpu_baseline',{}).get('value'), d['gpu_launches'])"

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/c58_pytest.txt; cat gpurun_out/c58_pytest.txt
for w in cfg2 cfg4 cfg1; do
timeout 900 python bench.py --steps 5 --warmup 3 --workload $w > gpurun_out/c58_bench_$w.json 2> gpurun_out/c58_bench_$w.err
python -c "
import json;d=json.loads(open('gpurun_out/c58_bench_$w.json').read().strip().split('\n')[-1]);print('$w', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'alu', round(d['roofline']['alu_pipe']['frac'],3), 'cpu', d.get('cpu_baseline',{}).get('value'))"
done

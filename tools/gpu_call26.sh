#!/bin/bash
# round-2 checkpoint: every GPU test, smoke, one bench line per BASELINE configuration, the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c26_pytest.txt
cat gpurun_out/c26_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c26_smoke.txt 2>&1; cat gpurun_out/c26_smoke.txt
for w in cfg2 cfg1 cfg4 cfg5; do
  timeout 900 python bench.py --workload $w > gpurun_out/c26_bench_$w.json 2> gpurun_out/c26_bench_$w.err
  python -c "
import json,sys;d=json.loads(open('gpurun_out/c26_bench_$w.json').read().strip().split('\n')[-1]);print('$w', round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'cpu', round(d['cpu_baseline']['value'],3))"
done
timeout 900 python bench.py --workload cfg3shard --steps 2 --warmup 3 > gpurun_out/c26_bench_cfg3shard.json 2> gpurun_out/c26_bench_cfg3shard.err
python -c "
import json,sys;d=json.loads(open('gpurun_out/c26_bench_cfg3shard.json').read().strip().split('\n')[-1]);print('cfg3shard', round(d['value'],1), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c26_bench_reference.json 2> gpurun_out/c26_bench_reference.err
tail -c 600 gpurun_out/c26_bench_reference.json
echo done

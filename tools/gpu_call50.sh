#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 0 1; do
  if [ $v = 1 ]; then export BSA_MAPPED_OUT=1; else unset BSA_MAPPED_OUT; fi
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().split('\n')[-1]);print('mapped=$v cfg2 dev', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2))"
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload cfg4 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().split('\n')[-1]);print('mapped=$v cfg4 dev', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2))"
done 2>&1 | tee gpurun_out/c50_mapped_out.txt

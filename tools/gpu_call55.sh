#!/bin/bash
python tools/quick_bench.py 1000 4 | tail -2
python tools/quick_bench.py 300 4 | tail -2
python bench.py --steps 10 --warmup 3 --workload cfg1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().split('\n')[-1]);print('cfg1 dev', round(d['value'],1), round(d['ms_per_step'],2), 'wall', round(d['config']['wall_ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2))"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2

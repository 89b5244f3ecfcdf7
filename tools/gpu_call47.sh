#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck python tools/sanitizer_probe.py 2>&1 | grep -v "^=========\s*$" | tail -14 > gpurun_out/c47_racecheck.txt
cat gpurun_out/c47_racecheck.txt
timeout 600 python bench.py --steps 5 --warmup 3 --workload cfg5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().split('\n')[-1]);print('cfg5', d['value'], d['ms_per_step'])"
timeout 600 python -m pytest tests/test_gpu_zconfigs_at_size.py -m gpu -x -q -k cfg5 2>&1 | tail -2

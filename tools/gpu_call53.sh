#!/bin/bash
for i in 1 2 3; do python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-90; done
BSA_MAPPED_OUT=0 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-90
timeout 600 python bench.py --steps 20 --warmup 5 --workload cfg5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().split('\n')[-1]);print('cfg5', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"

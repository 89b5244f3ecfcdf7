#!/bin/bash
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --gpus 2 --steps 3 --warmup 3 --single-process --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().split('\n')[-1]);print('single-process N=2 cfg2', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1)); open('gpurun_out/c61_bench_cfg2_sp_n2.json','w').write(json.dumps(d))"

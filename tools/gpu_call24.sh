#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi_device.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --single-process > gpurun_out/c24_bench_cfg2_sp.json 2> gpurun_out/c24_sp.err
python -c "
import json;d=json.loads(open('gpurun_out/c24_bench_cfg2_sp.json').read().strip().split('\n')[-1]);print('single-process', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"

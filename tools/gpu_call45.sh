#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --single-process --no-cpu-baseline > gpurun_out/c45_bench_cfg2_sp_n1.json 2> gpurun_out/c45_sp.err
python -c "
import json;d=json.loads(open('gpurun_out/c45_bench_cfg2_sp_n1.json').read().strip().split('\n')[-1]);print('single-process N=1 cfg2', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['n_sequences'])"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3

#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_multi_device.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --gpus 2 --steps 3 --warmup 3 --single-process --no-cpu-baseline > gpurun_out/c28_bench_cfg2_sp_n2.json 2> gpurun_out/c28_sp.err
python -c "
import json;d=json.loads(open('gpurun_out/c28_bench_cfg2_sp_n2.json').read().strip().split('\n')[-1]);print('single-process N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['n_sequences'])"
tail -2 gpurun_out/c28_sp.err

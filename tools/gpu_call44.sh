#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_parity.py::test_alphabet_limit_and_odd_even_lengths -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --gpus 8 --steps 3 --warmup 3 --single-process --no-cpu-baseline > gpurun_out/c44_bench_cfg2_sp_n8.json 2> gpurun_out/c44_sp.err
python -c "
import json;d=json.loads(open('gpurun_out/c44_bench_cfg2_sp_n8.json').read().strip().split('\n')[-1]);print('single-process N=8 cfg2', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['n_sequences'])"
timeout 600 python bench.py --gpus 8 --steps 2 --warmup 3 --single-process --no-cpu-baseline --workload cfg4 > gpurun_out/c44_bench_cfg4_sp_n8.json 2> gpurun_out/c44_sp4.err
python -c "
import json;d=json.loads(open('gpurun_out/c44_bench_cfg4_sp_n8.json').read().strip().split('\n')[-1]);print('single-process N=8 cfg4', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['n_sequences'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/cfg3_run.py 2> gpurun_out/c44_cfg3.err | tail -1 > gpurun_out/c44_cfg3_100k_8gpu.json
cut -c1-600 gpurun_out/c44_cfg3_100k_8gpu.json

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_zconfigs_at_size.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c6_pytest_configs.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c6_smoke.txt 2>&1
for w in cfg2 cfg1 cfg4 cfg5; do
  python bench.py --workload $w > gpurun_out/c6_bench_$w.json 2> gpurun_out/c6_bench_$w.err
done
python bench.py --workload cfg3shard --steps 2 --warmup 3 > gpurun_out/c6_bench_cfg3shard.json 2> gpurun_out/c6_bench_cfg3shard.err
python bench.py --single-process --gpus 1 --workload cfg2 --no-cpu-baseline > gpurun_out/c6_bench_cfg2_single_process.json 2> gpurun_out/c6_bench_cfg2_sp.err
nproc > gpurun_out/c6_nproc.txt; free -g >> gpurun_out/c6_nproc.txt
echo done

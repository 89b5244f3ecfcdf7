#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c7_pytest.txt
for round in 1 2; do
  for v in cur wave1row; do
    echo "[$v]"
    if [ "$v" = cur ]; then BSA_CFG5_NOCHECK=1 python tools/cfg5_run.py 2>&1 | tail -1; else BSA_CFG5_NOCHECK=1 BSA_LIB_PATH=$PWD/tools/microbench/libbsa_$v.so python tools/cfg5_run.py 2>&1 | tail -1; fi
  done
done > gpurun_out/c7_ab_wave.txt
BSA_CFG5_NOCHECK=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c7_launches_cfg5.csv python tools/cfg5_run.py > /dev/null 2>&1
python bench.py --single-process --gpus 1 --workload cfg2 --no-cpu-baseline > gpurun_out/c7_bench_cfg2_single_process.json 2> gpurun_out/c7_bench_cfg2_sp.err
echo done

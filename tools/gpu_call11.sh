#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs_at_size.py -m gpu -x -q -k "wave or path or titin or long or cfg5 or traceback" 2>&1 | tail -8 > gpurun_out/c11_pytest.txt
for round in 1 2; do
  for v in cur wave2rows; do
    echo "[$v]"
    if [ "$v" = cur ]; then BSA_CFG5_NOCHECK=1 python tools/cfg5_run.py 2>&1 | tail -1; else BSA_CFG5_NOCHECK=1 BSA_LIB_PATH=$PWD/tools/microbench/libbsa_$v.so python tools/cfg5_run.py 2>&1 | tail -1; fi
  done
done > gpurun_out/c11_cfg5.txt
BSA_CFG5_NOCHECK=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c11_launches_cfg5.csv python tools/cfg5_run.py > /dev/null 2>&1
echo done

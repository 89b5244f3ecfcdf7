#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/c48_pytest.txt; cat gpurun_out/c48_pytest.txt
tools/ab_run.sh "python tools/quick_bench.py 10000 2" cur nopadslot > gpurun_out/c48_ab_padslot.txt 2>&1
for v in cur nopadslot cur nopadslot; do
  if [ "$v" = cur ]; then python tools/quick_ovm.py 1000 50000 2>&1 | head -1 | sed "s/^/[$v] /"; else BSA_LIB_PATH=$PWD/tools/microbench/libbsa_$v.so python tools/quick_ovm.py 1000 50000 2>&1 | head -1 | sed "s/^/[$v] /"; fi
done >> gpurun_out/c48_ab_padslot.txt 2>&1
cat gpurun_out/c48_ab_padslot.txt

#!/bin/bash
mkdir -p gpurun_out
for n in 10000 1000; do tools/ab_run.sh "python tools/quick_bench.py $n 2" cur sk6 sk10; done > gpurun_out/c43_ab_smallk.txt 2>&1
cat gpurun_out/c43_ab_smallk.txt
for v in cur sk6 sk10 cur sk6 sk10; do
  if [ "$v" = cur ]; then python tools/quick_ovm.py 1000 50000 2>&1 | head -1 | sed "s/^/[$v] /"; else BSA_LIB_PATH=$PWD/tools/microbench/libbsa_$v.so python tools/quick_ovm.py 1000 50000 2>&1 | head -1 | sed "s/^/[$v] /"; fi
done

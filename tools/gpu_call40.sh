#!/bin/bash
mkdir -p gpurun_out
for n in 1000 3000 10000; do tools/ab_run.sh "python tools/quick_bench.py $n 2" cur st8 st16; done > gpurun_out/c40_ab_streams.txt 2>&1
cat gpurun_out/c40_ab_streams.txt
tools/ab_run.sh "python tools/quick_ovm.py 1000 50000" cur st8 | grep -v "^$"

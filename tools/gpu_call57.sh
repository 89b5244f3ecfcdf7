#!/bin/bash
mkdir -p gpurun_out
{
tools/ab_run.sh "python tools/quick_bench.py 10000 2" cur cs384 cs1024 w10e
for t in 20000 80000 20000 80000; do echo "[items $t] $(BSA_TARGET_ITEMS=$t python tools/quick_bench.py 10000 2 | tail -1)"; done
} > gpurun_out/c57_ab_misc.txt 2>&1
cat gpurun_out/c57_ab_misc.txt

#!/bin/bash
mkdir -p gpurun_out
tools/ab_run.sh "python tools/quick_bench.py 10000 2" cur w10 u4 u1 cb8k > gpurun_out/c32_ab_variants_cfg2.txt 2>&1
cat gpurun_out/c32_ab_variants_cfg2.txt

#!/bin/bash
mkdir -p gpurun_out
tools/ab_run.sh "python tools/quick_bench.py 10000 2" cur tail2 tail1 tail4 > gpurun_out/c56_ab_tail.txt 2>&1
cat gpurun_out/c56_ab_tail.txt

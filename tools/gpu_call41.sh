#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/c41_pytest.txt; cat gpurun_out/c41_pytest.txt
for n in 10000 1000; do tools/ab_run.sh "python tools/quick_bench.py $n 2" cur noetag; done > gpurun_out/c41_ab_etag.txt 2>&1
cat gpurun_out/c41_ab_etag.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hclust.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python tools/bench_hclust.py 10000 > gpurun_out/c27_hclust_10k.json 2>&1; cat gpurun_out/c27_hclust_10k.json | cut -c1-260
timeout 900 python tools/bench_hclust_large.py 50000 single average > gpurun_out/c27_hclust_50k.json 2> gpurun_out/c27_hclust_50k.err; cat gpurun_out/c27_hclust_50k.json | cut -c1-260

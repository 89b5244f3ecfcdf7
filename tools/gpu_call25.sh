#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/bench_hclust_large.py 50000 single average > gpurun_out/c25_hclust_50k.json 2> gpurun_out/c25_hclust_50k.err
cat gpurun_out/c25_hclust_50k.json; tail -2 gpurun_out/c25_hclust_50k.err
timeout 1200 python tools/bench_hclust_large.py 100000 single > gpurun_out/c25_hclust_100k.json 2> gpurun_out/c25_hclust_100k.err
cat gpurun_out/c25_hclust_100k.json; tail -2 gpurun_out/c25_hclust_100k.err

#!/bin/bash
mkdir -p gpurun_out
# K1s in the frame over the aligned stream: parity of the score-only paths, then cfg4-shape throughput (unaligned = before)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c33_pytest.txt
cat gpurun_out/c33_pytest.txt
tools/ab_run.sh "python tools/quick_ovm.py 1000 50000" cur unaligned > gpurun_out/c33_ab_ovm.txt 2>&1
cat gpurun_out/c33_ab_ovm.txt
tools/ab_run.sh "python tools/quick_bench.py 10000 2" u4w10 u8 u8w10 u4w12 > gpurun_out/c33_ab_variants_cfg2.txt 2>&1
cat gpurun_out/c33_ab_variants_cfg2.txt

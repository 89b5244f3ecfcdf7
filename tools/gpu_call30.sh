#!/bin/bash
# aligned two-row stream: parity first (the parity tests that exercise score + identity), then same-box A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c30_pytest.txt
cat gpurun_out/c30_pytest.txt
tools/ab_run.sh "python tools/quick_bench.py 10000 2" cur unaligned > gpurun_out/c30_ab_aligned_cfg2.txt 2>&1
cat gpurun_out/c30_ab_aligned_cfg2.txt
BSA_PROFILE_GROUPS=1 python tools/quick_bench.py 10000 1 > gpurun_out/c30_groups.txt 2>&1

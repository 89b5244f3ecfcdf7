#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c5_pytest.txt
python tools/quick_bench.py 10000 2 > gpurun_out/c5_quick.txt 2>&1
BSA_PROFILE_GROUPS=1 python tools/quick_bench.py 10000 0 > gpurun_out/c5_groups.txt 2>&1
echo done

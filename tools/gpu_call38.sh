#!/bin/bash
mkdir -p gpurun_out
BSA_PROFILE_GROUPS=1 python tools/quick_bench.py 1000 1 > gpurun_out/c38_groups_cfg1.txt 2>&1
python tools/quick_bench.py 1000 3 | tail -2
python tools/quick_bench.py 3000 3 | tail -1

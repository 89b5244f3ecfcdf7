#!/bin/bash
mkdir -p gpurun_out
for m in 1048576 4194304 16777216; do echo "min item $m"; BSA_MIN_ITEM_CELLS=$m python tools/quick_bench.py 1000 3 | tail -1; BSA_MIN_ITEM_CELLS=$m python tools/quick_bench.py 3000 2 | tail -1; done
python tools/quick_bench.py 10000 2 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/c65_pytest.txt; cat gpurun_out/c65_pytest.txt
python tools/quick_ovm.py 1000 50000 2>&1 | head -1

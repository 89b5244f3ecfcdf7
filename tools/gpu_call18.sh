#!/bin/bash
mkdir -p gpurun_out
BSA_CFG5_REPS=6 BSA_CFG5_NOCHECK=1 tools/ab_run.sh "python tools/cfg5_run.py" cur afma nohi > gpurun_out/c18_ab.txt 2>&1
grep -o '^\[[a-z]*\]\|"kernel_ms_all_reps": [^]]*]' gpurun_out/c18_ab.txt | paste - -
BSA_LIB_PATH=$PWD/tools/microbench/libbsa_afma.so python tools/wave_probe.py 34350 35000 1
BSA_LIB_PATH=$PWD/tools/microbench/libbsa_afma.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs_at_size.py -m gpu -x -q -k "wave or titin or cfg5" 2>&1 | tail -3

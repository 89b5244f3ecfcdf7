#!/bin/bash
# Round-2, first GPU call: baseline tests, per-group breakdown of cfg2, and the A/B runs prepared in round 1.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c1_pytest.txt
BSA_PROFILE_GROUPS=1 python tools/quick_bench.py 10000 1 > gpurun_out/c1_groups.txt 2>&1
run() { # name cmd...
  local v=$1; shift
  if [ "$v" = cur ]; then "$@" 2>&1 | tail -4; else BSA_LIB_PATH=$PWD/tools/microbench/libbsa_$v.so "$@" 2>&1 | tail -4; fi
}
for round in 1 2; do
  for v in cur floop; do echo "[$v]"; run $v python tools/quick_bench.py 10000 2; done
done > gpurun_out/c1_ab_floop_cfg2.txt
for round in 1 2; do
  for v in cur floop; do echo "[$v]"; run $v python tools/quick_ovm.py 1000 50000; done
done > gpurun_out/c1_ab_floop_ovm.txt
export BSA_CFG5_NOCHECK=1
for round in 1 2; do
  for v in cur b16 pf24 b16pf12; do echo "[$v]"; run $v python tools/cfg5_run.py; done
done > gpurun_out/c1_ab_wave.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt
echo done

#!/bin/bash
mkdir -p gpurun_out
{
echo "compute-sanitizer over tools/sanitizer_probe.py (every kernel incl. the aligned-stream two-row blocks and the 16-bit frame kernels), final build of round 2:"
for t in memcheck racecheck synccheck; do
  echo "--- $t"
  timeout 900 compute-sanitizer --tool $t python tools/sanitizer_probe.py 2>&1 | grep -v "^=========\s*$" | tail -6
done
} > gpurun_out/c46_sanitizer.txt 2>&1
cat gpurun_out/c46_sanitizer.txt

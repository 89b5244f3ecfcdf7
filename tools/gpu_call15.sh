#!/bin/bash
mkdir -p gpurun_out
{
python tools/wave_probe.py 34350 8704 1
python tools/wave_probe.py 34350 8704 4
python tools/wave_probe.py 34350 35000 1
python tools/wave_probe.py 34350 35000 4
} > gpurun_out/c15_probe.txt 2>&1
cat gpurun_out/c15_probe.txt

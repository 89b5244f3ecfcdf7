#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c66_smoke.txt 2>&1; tail -1 gpurun_out/c66_smoke.txt

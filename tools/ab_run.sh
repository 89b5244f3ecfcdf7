#!/bin/bash
# usage (on the GPU box): tools/ab_run.sh "<python command>" name1 name2 ...   ("cur" = the in-tree library)
# runs every variant twice, interleaved, and prints the last line of each run
cmd=$1; shift
for round in 1 2; do
  for v in "$@"; do
    if [ "$v" = cur ]; then out=$($cmd 2>&1 | tail -1); else out=$(BSA_LIB_PATH=$PWD/tools/microbench/libbsa_$v.so $cmd 2>&1 | tail -1); fi
    echo "[$v] $out"
  done
done

#!/bin/bash
# Build variants of libbioshell_align.so for same-box A/B timing: tools/ab_build.sh name "-DBSA_X=.." ...
# Output: tools/microbench/libbsa_<name>.so (git-ignored, travels with gpurun). Use with BSA_LIB_PATH.
set -e
cd "$(dirname "$0")/../bioshell_b200/csrc"
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -cudart static "$@" \
     -o ../../tools/microbench/libbsa_$name.so bsa_api.cu
echo "built $name"

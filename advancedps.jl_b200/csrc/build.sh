#!/bin/bash
# Builds libaps_b200.so for sm_100a. -fmad=false: no implicit contraction, so the shared
# arithmetic (include/aps_math.h) matches the gcc -ffp-contract=off build of the oracle bit for bit.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off,-mfma -Xptxas -v \
  -I../../include -shared -o ${APS_OUT:-../libaps_b200.so} aps_api.cu -lcudart "$@"

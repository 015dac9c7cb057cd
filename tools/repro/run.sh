#!/bin/bash
# Build the three variants of the division repro and run them on the GPU (under gpurun).
cd "$(dirname "$0")"
F="-std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr"
nvcc -O3 $F div_miscompile.cu -o /tmp/div_O3 && nvcc -O3 -Xptxas -O0 $F div_miscompile.cu -o /tmp/div_O0 && nvcc -O3 -DLPC_DIV_OPAQUE_NEG $F div_miscompile.cu -o /tmp/div_O3_opaque || exit 3
for v in O3 O0 O3_opaque; do echo "== ptxas variant $v"; /tmp/div_$v ${1:-6}; echo "exit $?"; done

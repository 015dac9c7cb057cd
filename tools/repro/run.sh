#!/bin/bash
# Build the variants of the division repro and run them on the GPU (under gpurun).
cd "$(dirname "$0")"
F="-std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -diag-suppress 821"
build() { nvcc -O3 $F "${@:2}" div_miscompile.cu -o /tmp/div_$1 || exit 3; }
build O3
build O0 -Xptxas -O0
build O1 -Xptxas -O1
build O3_opaque -DLPC_DIV_OPAQUE_NEG
build O3_fix1 -DLPC_DIV_FIX=1
build O3_fix2 -DLPC_DIV_FIX=2
build cicc0_ptxas3 -Xcicc -O0
build O3_fix3 -DLPC_DIV_FIX=3
build O3_fix3_r64 -DLPC_DIV_FIX=3 -maxrregcount=64
for v in O3 O0 cicc0_ptxas3 O3_fix3 O3_fix3_r64; do echo "== variant $v"; /tmp/div_$v ${1:-6}; echo "exit $?"; done

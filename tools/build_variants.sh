#!/bin/bash
# Builds libgramhead variants with different backward ring depths (A stages, B stages, store buffers) for sweeps on the GPU box.
cd "$(dirname "$0")/../heuristique_style_transfer_code_b200/csrc" || exit 1
mkdir -p variants
for v in "$@"; do
  IFS=_ read -r a b s <<< "$v"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
       -DGH_BP_A_STAGES=$a -DGH_BP_B_STAGES=$b -DGH_BP_STORE_BUFS=$s gramhead.cu -o variants/libgramhead_${v}.so || echo "variant $v failed"
done
ls -la variants

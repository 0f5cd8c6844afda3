#!/bin/bash
# bring-up of new kernels first (isolated processes), then the full round
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r}
timeout 1200 python tools/gpu_bringup.py gemm_layouts bwd_pool bwd_dense attn_fwd_bwd generic_pool module_parity > $OUT/${TAG}_bringup.log 2>&1
echo "bringup exit=$?"; grep -E "rel_err|rel=|PASS|FAIL|SUMMARY|error|Error" $OUT/${TAG}_bringup.log | head -120
bash tools/gpu_round.sh $TAG ${2:-}

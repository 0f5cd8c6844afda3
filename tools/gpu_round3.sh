#!/bin/bash
# Bring-up of the CTA-pair kernels (isolated processes), then - only if they pass - the GPU test suite and timings.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 1500 python tools/gpu_bringup.py pair_fwd_min pair_bwd_min pair_fwd pair_fwd_dense pair_bwd pair_bwd_dense > $OUT/${TAG}_bringup.log 2>&1
rc=$?
echo "bringup exit=$rc"; grep -E "pair-|PASS|FAIL|SUMMARY|rror|device error" $OUT/${TAG}_bringup.log | head -150
timeout 900 python tools/gpu_bringup.py timing_pair > $OUT/${TAG}_timing_pair.log 2>&1
echo "timing exit=$?"; grep -E "^fwd|^bwd|PASS|FAIL" $OUT/${TAG}_timing_pair.log | head -80
if [ "$rc" = "0" ]; then
  echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -15 $OUT/${TAG}_pytest_gpu.log
fi

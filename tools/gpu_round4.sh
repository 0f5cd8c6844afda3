#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r}
timeout 900 python tools/gpu_bringup.py pair_bwd_min pair_bwd pair_bwd_dense pair_fwd_dense > $OUT/${TAG}_bringup.log 2>&1
echo "bringup exit=$?"; grep -E "pair-|PASS|FAIL|SUMMARY|rror|device error" $OUT/${TAG}_bringup.log | head -80
timeout 600 python tools/gpu_bringup.py timing_pair > $OUT/${TAG}_timing_pair.log 2>&1
echo "timing exit=$?"; grep -E "^fwd|^bwd|PASS|FAIL" $OUT/${TAG}_timing_pair.log | head -80
# ncu: each capture right after the same command exited 0 without ncu
prof() {  # name, args...
  local name=$1; shift
  python tools/prof_one.py "$@" > $OUT/${TAG}_plain_${name}.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 2 -c 1 -o $OUT/${TAG}_${name} \
      python tools/prof_one.py "$@" > $OUT/${TAG}_ncu_${name}.log 2>&1
  echo "ncu $name exit=$?"
}
prof fwd_s1_f32 fwd 256 3136 256 f32 1
prof fwd_s3_f32 fwd 1024 196 256 f32 1
prof fwd_s2_bf16 fwd 512 784 256 bf16 1
prof bwd_s1_f32 bwd 256 3136 256 f32 1
prof bwd_s2_f32 bwd 512 784 256 f32 1
ls -la $OUT | grep ${TAG}

#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r}
timeout 900 python tools/gpu_bringup.py pair_bwd_min pair_bwd pair_bwd_dense > $OUT/${TAG}_bringup.log 2>&1
rc=$?
echo "bringup exit=$rc"; grep -E "pair-|PASS|FAIL|SUMMARY|rror|device error" $OUT/${TAG}_bringup.log | head -60
timeout 600 python tools/gpu_bringup.py timing_pair > $OUT/${TAG}_timing_pair.log 2>&1
echo "timing exit=$?"; grep -E "^bwd|^fwd.*pair=1|PASS|FAIL" $OUT/${TAG}_timing_pair.log | head -80
if [ "$rc" = "0" ]; then
  echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -15 $OUT/${TAG}_pytest_gpu.log
fi
prof() {  # name, args...
  local name=$1; shift
  python tools/prof_one.py "$@" > $OUT/${TAG}_plain_${name}.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 2 -c 1 -o $OUT/${TAG}_${name} \
      python tools/prof_one.py "$@" > $OUT/${TAG}_ncu_${name}.log 2>&1
  echo "ncu $name exit=$?"
  ncu -i $OUT/${TAG}_${name}.ncu-rep --page raw --csv > $OUT/${TAG}_${name}_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_${name}.ncu-rep --page source --csv > $OUT/${TAG}_${name}_source.csv 2>/dev/null
  ncu -i $OUT/${TAG}_${name}.ncu-rep --page details > $OUT/${TAG}_${name}_details.txt 2>/dev/null
  rm -f $OUT/${TAG}_${name}.ncu-rep
}
prof bwd_s1_f32 bwd 256 3136 256 f32 1
prof bwd_s3_f32 bwd 1024 196 256 f32 1

#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r}
timeout 900 python tools/gpu_bringup.py pair_bwd_min pair_bwd pair_bwd_dense > $OUT/${TAG}_bringup.log 2>&1
rc=$?
echo "bringup exit=$rc"; grep -E "PASS|FAIL|SUMMARY|rror|device error" $OUT/${TAG}_bringup.log | head -30
timeout 600 python tools/gpu_bringup.py timing_pair > $OUT/${TAG}_timing_pair.log 2>&1
echo "timing exit=$?"; grep -E "^bwd.*pair=1|PASS|FAIL" $OUT/${TAG}_timing_pair.log | head -40
[ "$rc" = "0" ] || exit 1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_module.py -x -q -m gpu -k "backward or prefetch or golden or module" > $OUT/${TAG}_pytest_bwd.log 2>&1; echo "pytest exit=$?"; tail -4 $OUT/${TAG}_pytest_bwd.log
timeout 900 python tools/bench_sweep.py > $OUT/${TAG}_sweep.md 2> $OUT/${TAG}_sweep.err; echo "sweep exit=$?"; wc -l $OUT/${TAG}_sweep.md; tail -3 $OUT/${TAG}_sweep.err
prof() {
  local name=$1; shift
  local rx=$1; shift
  python tools/prof_one.py "$@" > $OUT/${TAG}_plain_${name}.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o $OUT/${TAG}_${name} \
      python tools/prof_one.py "$@" > $OUT/${TAG}_ncu_${name}.log 2>&1
  echo "ncu $name exit=$?"
  ncu -i $OUT/${TAG}_${name}.ncu-rep --page raw --csv > $OUT/${TAG}_${name}_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_${name}.ncu-rep --page source --csv > $OUT/${TAG}_${name}_source.csv 2>/dev/null
  ncu -i $OUT/${TAG}_${name}.ncu-rep --page details > $OUT/${TAG}_${name}_details.txt 2>/dev/null
  rm -f $OUT/${TAG}_${name}.ncu-rep
}
prof bwd_s2_f32 gram_bwd bwd 512 784 256 f32 -1
prof bwd_s3_f32 gram_bwd bwd 1024 196 256 f32 -1

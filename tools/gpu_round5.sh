#!/bin/bash
# ncu captures of the pair kernels, exported to CSV on the box (the .ncu-rep files are too large to bring back together)
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r}
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
prof bwd_s1_bf16 bwd 256 3136 256 bf16 1
prof fwd_s1_f32 fwd 256 3136 256 f32 1
prof fwd_s2_bf16 fwd 512 784 256 bf16 1
ls -la $OUT | grep ${TAG}

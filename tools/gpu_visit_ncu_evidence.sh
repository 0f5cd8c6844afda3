#!/bin/bash
# Evidence refresh for the shipped default (channels_last backbone -> NHWC CTA-pair kernels): bench, ncu launch list of
# the same bench command, and one `--set full` capture per Gram kernel/stage, exported to text on the box.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== bench"; timeout 1500 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit=$?"; head -c 1200 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
echo "== ncu launch list (inference step)"
python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu --skip-handoff > $OUT/${TAG}_ncu_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu --skip-handoff > $OUT/${TAG}_ncu_list.log 2>&1
echo "ncu list exit=$?"
prof() {  # kind C HW dtype layout
  local name=$1_$2_$3_$4_$5
  python tools/prof_one.py $1 $2 $3 256 $4 -1 $5 > $OUT/${TAG}_plain_${name}.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_$1 -s 2 -c 1 -o $OUT/${TAG}_${name} \
      python tools/prof_one.py $1 $2 $3 256 $4 -1 $5 > $OUT/${TAG}_ncu_${name}.log 2>&1
  echo "ncu $name exit=$?"
  ncu -i $OUT/${TAG}_${name}.ncu-rep --page raw --csv > $OUT/${TAG}_${name}_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_${name}.ncu-rep --page details > $OUT/${TAG}_${name}_details.txt 2>/dev/null
  rm -f $OUT/${TAG}_${name}.ncu-rep $OUT/${TAG}_plain_${name}.log
}
for layout in nhwc; do
  prof fwd 256 3136 f32 $layout
  prof fwd 512 784 f32 $layout
  prof fwd 1024 196 f32 $layout
  prof bwd 256 3136 f32 $layout
  prof bwd 512 784 f32 $layout
  prof bwd 1024 196 f32 $layout
done
prof fwd 256 3136 f32 nchw
ls -la $OUT | grep ${TAG} | head -60

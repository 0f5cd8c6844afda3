#!/bin/bash
# usage: tools/gpu_prof.sh TAG name regex prof_one-args...   (one ncu --set full capture exported to text)
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=$1; name=$2; rx=$3; shift 3
python tools/prof_one.py "$@" > $OUT/${TAG}_plain_${name}.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o $OUT/${TAG}_${name} \
    python tools/prof_one.py "$@" > $OUT/${TAG}_ncu_${name}.log 2>&1
echo "ncu $name exit=$?"
ncu -i $OUT/${TAG}_${name}.ncu-rep --page raw --csv > $OUT/${TAG}_${name}_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_${name}.ncu-rep --page source --csv > $OUT/${TAG}_${name}_source.csv 2>/dev/null
ncu -i $OUT/${TAG}_${name}.ncu-rep --page details > $OUT/${TAG}_${name}_details.txt 2>/dev/null
rm -f $OUT/${TAG}_${name}.ncu-rep

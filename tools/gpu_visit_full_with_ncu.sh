#!/bin/bash
# Full visit: pair-kernel bring-up, timings, GPU tests, smoke, bench (+reference arm), ncu launch list and full captures
# exported to text on the box (the .ncu-rep files together exceed what gpurun brings back).
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $OUT/${TAG}_gpu.txt 2>&1
timeout 900 python tests/tools/gpu_bringup.py pair_fwd_min pair_bwd_min pair_bwd pair_bwd_dense > $OUT/${TAG}_bringup.log 2>&1
rc=$?
echo "bringup exit=$rc"; grep -E "pair-|PASS|FAIL|SUMMARY|rror|device error" $OUT/${TAG}_bringup.log | head -60
timeout 600 python tests/tools/gpu_bringup.py timing_pair > $OUT/${TAG}_timing_pair.log 2>&1
echo "timing exit=$?"; grep -E "^bwd.*pair=1|^fwd.*pair=1|PASS|FAIL" $OUT/${TAG}_timing_pair.log | head -80
[ "$rc" = "0" ] || exit 1
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -8 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit=$?"; tail -3 $OUT/${TAG}_smoke.log
echo "== bench"; timeout 1500 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit=$?"; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>/dev/null; cat $OUT/${TAG}_bench_ref.json
if [ "${2:-}" != "skip-ncu" ]; then
  echo "== ncu launch list (inference step)"
  python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu > $OUT/${TAG}_ncu_plain.log 2>&1 && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu > $OUT/${TAG}_ncu_list.log 2>&1
  echo "ncu list exit=$?"
  prof() {  # name, kernel regex, args...
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
  prof fwd_s1_f32 gram_fwd fwd 256 3136 256 f32 -1
  prof fwd_s2_f32 gram_fwd fwd 512 784 256 f32 -1
  prof fwd_s3_f32 gram_fwd fwd 1024 196 256 f32 -1
  prof bwd_s1_f32 gram_bwd bwd 256 3136 256 f32 -1
  prof bwd_s2_f32 gram_bwd bwd 512 784 256 f32 -1
  prof bwd_s3_f32 gram_bwd bwd 1024 196 256 f32 -1
fi
ls -la $OUT | grep ${TAG} | head -60

#!/bin/bash
# One GPU-box visit: GPU tests, smoke, bench, kernel timings, then the ncu passes (each only after its plain run exited 0).
# usage: tools/gpu_round.sh <tag> [skip-ncu]
set -u
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -5 $OUT/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit=$?"; tail -3 $OUT/${TAG}_smoke.log
echo "== timing"; timeout 900 python tools/gpu_bringup.py timing > $OUT/${TAG}_timing.log 2>&1; echo "timing exit=$?"; grep -E "fwd|bwd|torch" $OUT/${TAG}_timing.log | head -60
echo "== bench"; timeout 1500 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit=$?"; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>/dev/null; cat $OUT/${TAG}_bench_ref.json
if [ "${2:-}" != "skip-ncu" ]; then
  echo "== ncu launch list"
  python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu > $OUT/${TAG}_ncu_plain.log 2>&1 && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu > $OUT/${TAG}_ncu_list.log 2>&1
  echo "ncu list exit=$?"
  echo "== ncu full (gram fwd)"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gram_fwd_kernel -s 9 -c 3 -o $OUT/${TAG}_gram_fwd \
      python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu > $OUT/${TAG}_ncu_full.log 2>&1
  echo "ncu full exit=$?"
  echo "== ncu full (train kernels: gram bwd, attention GEMMs)"
  python bench.py --steps 1 --warmup 3 --train-steps 1 --skip-cpu > $OUT/${TAG}_ncu_plain2.log 2>&1 && \
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"gram_bwd2_kernel|umma_gemm_kernel" -c 12 -o $OUT/${TAG}_train_kernels \
      python bench.py --steps 1 --warmup 3 --train-steps 1 --skip-cpu > $OUT/${TAG}_ncu_full2.log 2>&1
  echo "ncu full2 exit=$?"; ls -la $OUT | tail -5
fi

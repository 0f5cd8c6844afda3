#!/bin/bash
# Round-end style validation: GPU suite (with durations), smoke, bench (both arms), camera + PatchGAN side benches, and the
# ncu launch list of the inference step of the same bench command.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -x -q -m gpu --durations=25 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit=$?"; grep -E "passed|failed|error" $OUT/${TAG}_pytest_gpu.log | tail -3
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit=$?"; tail -2 $OUT/${TAG}_smoke.log
echo "== bench"; timeout 1500 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit=$?"; python -c "
import json; d=json.load(open('$OUT/${TAG}_bench.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','cpu_baseline')}); print(d['roofline']); print(d['train']['value'], d['train']['ms_per_step'], d['train']['head'], d['train'].get('bf16_handoff')); print({k: v for k, v in list(d.items())[-14:]})"; tail -3 $OUT/${TAG}_bench.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_ref.json 2>/dev/null; cut -c1-300 $OUT/${TAG}_bench_ref.json
echo "== camera"; timeout 600 python tools/bench_camera.py > $OUT/${TAG}_camera.json 2> $OUT/${TAG}_camera.err; echo "camera exit=$?"; cut -c1-600 $OUT/${TAG}_camera.json
echo "== ncu launch list (inference step)"
python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu --skip-handoff --skip-patchgan --skip-reference-gpu > $OUT/${TAG}_ncu_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-train --skip-cpu --skip-handoff --skip-patchgan --skip-reference-gpu > $OUT/${TAG}_ncu_list.log 2>&1
echo "ncu list exit=$?"
ls -la $OUT | grep ${TAG} | head -30

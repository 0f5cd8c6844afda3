#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r}
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -6 $OUT/${TAG}_pytest_gpu.log
timeout 600 python tools/bench_backbone_modes.py 256 > $OUT/${TAG}_backbone_modes.log 2>&1; echo "modes exit=$?"; tail -3 $OUT/${TAG}_backbone_modes.log | cut -c1-230
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit=$?"; tail -2 $OUT/${TAG}_smoke.log
echo "== bench"; timeout 1500 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit=$?"; python -c "
import json; d=json.load(open('$OUT/${TAG}_bench.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['roofline']); print(d['train']['value'], d['train']['ms_per_step'], d['train']['roofline'])"; tail -3 $OUT/${TAG}_bench.err

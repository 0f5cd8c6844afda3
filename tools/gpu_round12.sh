#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r}
timeout 1200 python -m pytest tests/test_gpu_module.py tests/test_gpu_integration.py -x -q -m gpu > $OUT/${TAG}_pytest_mod.log 2>&1; echo "pytest exit=$?"; tail -8 $OUT/${TAG}_pytest_mod.log
timeout 600 python tools/bench_style.py > $OUT/${TAG}_style.json 2> $OUT/${TAG}_style.err; echo "style exit=$?"; cat $OUT/${TAG}_style.json; tail -3 $OUT/${TAG}_style.err
timeout 900 python bench.py --skip-cpu > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit=$?"; python -c "
import json; d=json.load(open('$OUT/${TAG}_bench.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','backbone_handoff')}); print(d['roofline']); print(d['train']['value'], d['train']['roofline']['kernel'], d['train']['roofline']['frac'])"; tail -3 $OUT/${TAG}_bench.err

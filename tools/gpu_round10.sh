#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r}
timeout 900 python -m pytest tests/test_gpu_module.py -x -q -m gpu > $OUT/${TAG}_pytest_module.log 2>&1; echo "pytest module exit=$?"; tail -12 $OUT/${TAG}_pytest_module.log
timeout 600 python tools/bench_backbone_modes.py 256 > $OUT/${TAG}_backbone_modes.log 2>&1; echo "modes exit=$?"; tail -4 $OUT/${TAG}_backbone_modes.log
timeout 900 python tools/bench_sweep.py > $OUT/${TAG}_sweep.md 2> $OUT/${TAG}_sweep.err; echo "sweep exit=$?"; wc -l $OUT/${TAG}_sweep.md

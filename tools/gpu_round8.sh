#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r}
timeout 900 python -m pytest tests/test_gpu_streaming.py -x -q -m gpu > $OUT/${TAG}_pytest_streaming.log 2>&1; echo "pytest streaming exit=$?"; tail -25 $OUT/${TAG}_pytest_streaming.log
timeout 600 python tools/bench_camera.py --frames 200 > $OUT/${TAG}_camera.json 2> $OUT/${TAG}_camera.err; echo "camera exit=$?"; cat $OUT/${TAG}_camera.json; tail -5 $OUT/${TAG}_camera.err

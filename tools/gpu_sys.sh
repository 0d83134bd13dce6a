#!/bin/bash
# GPU check of the System rows: bash tools/gpu_sys.sh <tag> [pytest -k expression]
TAG=${1:-sys}; K=${2:-}
O=gpurun_out/$TAG
mkdir -p $O
if [ -n "$K" ]; then timeout 900 python -m pytest tests -m gpu -q -k "$K" > $O/pytest_gpu.log 2>&1; else timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; fi
echo "pytest exit $?" >> $O/pytest_gpu.log; grep -E "^E  .*(assert|Error)|passed|failed|^FAILED" $O/pytest_gpu.log | cut -c1-220 | tail -25
timeout 600 python tools/tune.py --sys-only > $O/tune_sys.log 2>&1; echo "tune exit $?" >> $O/tune_sys.log; tail -45 $O/tune_sys.log

#!/bin/bash
# Quick GPU check: bash tools/gpu_quick.sh <tag> [pytest -k expression]
TAG=${1:-quick}; K=${2:-}
O=gpurun_out/$TAG
mkdir -p $O
if [ -n "$K" ]; then python -m pytest tests -m gpu -q -k "$K" > $O/pytest_gpu.log 2>&1; else python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; fi
echo "pytest exit $?" >> $O/pytest_gpu.log; grep -E "^E  .*(assert|Error)|passed|failed|^FAILED" $O/pytest_gpu.log | cut -c1-220 | tail -25
python tools/tune.py --rls-only > $O/tune_rls.log 2>&1; tail -8 $O/tune_rls.log

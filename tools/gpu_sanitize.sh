#!/bin/bash
# compute-sanitizer on the smallest representative cases (one tool per gpurun call).
# bash tools/gpu_sanitize.sh <tag> <memcheck|racecheck|synccheck|initcheck>
TAG=${1:-san}; TOOL=${2:-memcheck}
O=gpurun_out/$TAG
mkdir -p $O
# the same command must pass without the tool first
python -m pytest tests -m gpu -q -x -k "ragged_sizes or every_output_mask or golden or rollout_cost_argmin or eight_byte or rls_exact" > $O/plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests -m gpu -q -x -k "ragged_sizes or every_output_mask or golden or rollout_cost_argmin or eight_byte or rls_exact" > $O/$TOOL.log 2>&1
echo "sanitizer exit $?" >> $O/$TOOL.log
tail -3 $O/plain.log; grep -E "ERROR SUMMARY|passed|failed|Invalid|Race|hazard" $O/$TOOL.log | tail -8; tail -2 $O/$TOOL.log

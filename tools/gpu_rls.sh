#!/bin/bash
# RLS kernels: parity tests, then the throughput rows for each variant.  bash tools/gpu_rls.sh <tag>
TAG=${1:-rls}; O=gpurun_out/$TAG; mkdir -p $O
python -m pytest tests -m gpu -q -k "rls or identification" > $O/pytest_rls.log 2>&1; echo "pytest exit $?" >> $O/pytest_rls.log
grep -E "^E  .*(assert|Error)|passed|failed|^FAILED|pytest exit" $O/pytest_rls.log | cut -c1-200 | tail -8
for v in 0 1; do
  echo "== BLF_CCM_TUNE_RLS_PIPE=$v" | tee -a $O/tune_rls.log
  BLF_CCM_TUNE_RLS_PIPE=$v python tools/tune.py --rls-only >> $O/tune_rls.log 2>&1
  tail -5 $O/tune_rls.log
done

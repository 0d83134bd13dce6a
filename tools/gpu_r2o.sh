#!/bin/bash
TAG=${1:-r2o}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sys.py -m gpu -q -x -k "warp_specialised or rollout" > $O/pytest_sys.log 2>&1; echo "pytest exit $?" >> $O/pytest_sys.log
grep -E "^E  .*|passed|failed|^FAILED|pytest exit" $O/pytest_sys.log | cut -c1-220 | tail -6
for i in 1 2; do
python tools/prof_rollout.py 0.01
python tools/prof_rollout.py 0.0
python tools/prof_rollout.py 0.01 1024
done > $O/prof.log 2>&1; cat $O/prof.log
HSWEEP_VARIANTS=34 timeout 300 python tools/rollout_hsweep.py > $O/hsweep.log 2>&1; cat $O/hsweep.log

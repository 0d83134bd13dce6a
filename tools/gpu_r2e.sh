#!/bin/bash
# Round 2, fifth GPU call: rollout kernel tests + sweep (fourth form), ncu of it.
TAG=${1:-r2e}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_sys.py tests/test_rls.py -m gpu -q -x > $O/pytest_sys.log 2>&1; echo "pytest exit $?" >> $O/pytest_sys.log
grep -E "^E  .*|passed|failed|^FAILED|pytest exit" $O/pytest_sys.log | cut -c1-220 | tail -12
timeout 900 python tools/rollout_sweep.py > $O/rollout_sweep.log 2>&1; cat $O/rollout_sweep.log
export BLF_CCM_TUNE_ROLLOUT_WS=23
python tools/prof_rollout.py 0.01 > $O/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ccm_rollout_ws4 -s 4 -c 1 -o $O/prof_ws4_rho \
    python tools/prof_rollout.py 0.01 > $O/ncu_ws4_rho.log 2>&1
cat $O/prof_plain.log; tail -2 $O/ncu_ws4_rho.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file $O/launches_rollout.csv python tools/prof_rollout.py 0.01 > $O/ncu_l.log 2>&1
grep -E "ccm_" $O/launches_rollout.csv | tail -4 | cut -c1-220
ls -la $O

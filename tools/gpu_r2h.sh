#!/bin/bash
# Round 2: fifth rollout form -- parity tests, sweep against the third form, fixed vs per-step cost, ncu.
TAG=${1:-r2h}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sys.py -m gpu -q -x -k "warp_specialised or rollout" > $O/pytest_sys.log 2>&1; echo "pytest exit $?" >> $O/pytest_sys.log
grep -E "^E  .*|passed|failed|^FAILED|pytest exit" $O/pytest_sys.log | cut -c1-220 | tail -12
SWEEP_ONLY=auto,ws3,ws5 timeout 600 python tools/rollout_sweep.py > $O/rollout_sweep.log 2>&1; cat $O/rollout_sweep.log
HSWEEP_VARIANTS=33,34,35 timeout 600 python tools/rollout_hsweep.py > $O/hsweep.log 2>&1; cat $O/hsweep.log
export BLF_CCM_TUNE_ROLLOUT_WS=34
python tools/prof_rollout.py 0.01 > $O/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ccm_rollout_ws5 -s 4 -c 1 -o $O/prof_ws5_rho \
    python tools/prof_rollout.py 0.01 > $O/ncu_ws5_rho.log 2>&1
cat $O/prof_plain.log; tail -2 $O/ncu_ws5_rho.log
ls -la $O

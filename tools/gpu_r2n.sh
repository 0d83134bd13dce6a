#!/bin/bash
# Round 2: late programmatic dependent launch of the fifth rollout form (A/B), the 2^28-state test on
# its own with skip reasons, ncu launch list of the bench's headline loop.
TAG=${1:-r2n}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sys.py -m gpu -q -x -k "warp_specialised or rollout" > $O/pytest_sys.log 2>&1; echo "pytest exit $?" >> $O/pytest_sys.log
grep -E "^E  .*|passed|failed|^FAILED|pytest exit" $O/pytest_sys.log | cut -c1-220 | tail -6
for pdl in 0 1; do
  echo "BLF_CCM_TUNE_NO_PDL=$pdl"
  BLF_CCM_TUNE_NO_PDL=$pdl python tools/prof_rollout.py 0.01
  BLF_CCM_TUNE_NO_PDL=$pdl python tools/prof_rollout.py 0.0
  BLF_CCM_TUNE_NO_PDL=$pdl python tools/prof_rollout.py 0.01 1024
  BLF_CCM_TUNE_NO_PDL=$pdl HSWEEP_VARIANTS=34 timeout 300 python tools/rollout_hsweep.py
done > $O/pdl_ab.log 2>&1; cat $O/pdl_ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rs -k "config5 or config4" --durations=4 > $O/pytest_big.log 2>&1; echo "pytest exit $?" >> $O/pytest_big.log
grep -E "^E  .*|passed|failed|SKIPPED|pytest exit|s call" $O/pytest_big.log | cut -c1-220 | tail -8
PROF="python bench.py --gpus 1 --steps 20 --warmup 5 --only-main"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ccm_ -c 60 --csv --log-file $O/launches_main.csv $PROF > $O/ncu_launches.log 2>&1
grep -c ccm_ $O/launches_main.csv; grep ccm_ $O/launches_main.csv | tail -4 | cut -c1-260
ls -la $O

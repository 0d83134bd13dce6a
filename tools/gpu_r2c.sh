#!/bin/bash
# Round 2, third GPU call: host-path stream/schedule sweep, FP64 latency, ncu of the rollout producer.
TAG=${1:-r2c}
O=gpurun_out/$TAG
mkdir -p $O
./bipedal_locomotion_framework_b200/lib/fp64_lat > $O/fp64_lat.log 2>&1; cat $O/fp64_lat.log
timeout 900 python tools/host_sweep.py > $O/host_sweep.log 2>&1; cat $O/host_sweep.log
python -m pytest tests/test_gpu_sys.py tests/test_gpu_parity.py -m gpu -q -x -k "generalized or host" > $O/pytest_sel.log 2>&1; tail -3 $O/pytest_sel.log
python tools/tune.py --gf-only > $O/tune_gf.log 2>&1; cat $O/tune_gf.log
export BLF_CCM_TUNE_ROLLOUT_WS=3
python tools/prof_rollout.py 0.01 > $O/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ccm_rollout_ws2 -s 4 -c 1 -o $O/prof_ws2_rho \
    python tools/prof_rollout.py 0.01 > $O/ncu_ws2_rho.log 2>&1
cat $O/prof_plain.log; tail -2 $O/ncu_ws2_rho.log
python tools/prof_rollout.py 0.0 > $O/prof_plain0.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ccm_rollout_ws2 -s 4 -c 1 -o $O/prof_ws2_rho0 \
    python tools/prof_rollout.py 0.0 > $O/ncu_ws2_rho0.log 2>&1
cat $O/prof_plain0.log; tail -2 $O/ncu_ws2_rho0.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file $O/launches_rollout.csv python tools/prof_rollout.py 0.01 > $O/ncu_l.log 2>&1
grep -E "ccm_" $O/launches_rollout.csv | tail -8 | cut -c1-200
ls -la $O

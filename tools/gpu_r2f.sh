#!/bin/bash
# Round 2, sixth GPU call: FP64 issue throughput, fixed vs per-step cost of the rollout forms, ncu of the third form.
TAG=${1:-r2f}
O=gpurun_out/$TAG
mkdir -p $O
./bipedal_locomotion_framework_b200/lib/fp64_lat > $O/fp64_lat.log 2>&1; cat $O/fp64_lat.log
HSWEEP_VARIANTS=${HSWEEP_VARIANTS:-13,3} timeout 600 python tools/rollout_hsweep.py > $O/hsweep.log 2>&1; cat $O/hsweep.log
export BLF_CCM_TUNE_ROLLOUT_WS=13
python tools/prof_rollout.py 0.01 > $O/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ccm_rollout_ws3 -s 4 -c 1 -o $O/prof_ws3_rho \
    python tools/prof_rollout.py 0.01 > $O/ncu_ws3_rho.log 2>&1
cat $O/prof_plain.log; tail -2 $O/ncu_ws3_rho.log
python tools/prof_rollout.py 0.01 1024 > $O/prof_plain_1024.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ccm_rollout_ws3 -s 4 -c 1 -o $O/prof_ws3_rho_1024 \
    python tools/prof_rollout.py 0.01 1024 > $O/ncu_ws3_rho_1024.log 2>&1
cat $O/prof_plain_1024.log
ls -la $O

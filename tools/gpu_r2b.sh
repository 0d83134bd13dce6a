#!/bin/bash
# Round 2, second GPU call: all GPU tests again (after the worker-pool fix), host-path sweep,
# expansion bandwidth, rollout kernel sweep.   bash tools/gpu_r2b.sh <tag>
TAG=${1:-r2b}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E "^E  .*(assert|Error)|passed|failed|^FAILED|pytest exit|^[0-9.]+s (call|setup)" $O/pytest_gpu.log | cut -c1-220 | tail -30
./bipedal_locomotion_framework_b200/lib/HostExpandUnitTests > $O/host_expand.log 2>&1
./bipedal_locomotion_framework_b200/lib/HostExpandUnitTests --bandwidth 16 >> $O/host_expand.log 2>&1
cat $O/host_expand.log
timeout 600 python tools/host_sweep.py > $O/host_sweep.log 2>&1; cat $O/host_sweep.log
BLF_CCM_TUNE_HOST_NOEXPAND=1 timeout 300 python tools/host_sweep.py > $O/host_sweep_noexpand.log 2>&1; grep "chunk   65536" $O/host_sweep_noexpand.log
timeout 600 python tools/rollout_sweep.py > $O/rollout_sweep.log 2>&1; cat $O/rollout_sweep.log
ls -la $O

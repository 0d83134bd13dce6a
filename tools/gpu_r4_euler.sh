#!/bin/bash
# Round 2, fourth session: the floating-base Euler step -- GPU tests, tile sweep, C++ test, one full ncu capture.
TAG=${1:-r4b}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dyn.py tests/test_cpp_facade.py -m gpu -q -x -rs --durations=5 > $O/pytest_dyn.log 2>&1; echo "pytest exit $?" >> $O/pytest_dyn.log
grep -E "^E  .*|passed|failed|^FAILED|SKIPPED|pytest exit" $O/pytest_dyn.log | cut -c1-220 | tail -12
./bipedal_locomotion_framework_b200/lib/IntegratorUnitTests > $O/cpp_IntegratorUnitTests.log 2>&1; echo "IntegratorUnitTests exit $?"; tail -2 $O/cpp_IntegratorUnitTests.log
for kb in 12 16 24 32 48; do
echo "== BLF_CCM_TUNE_FBD_TILE_KB=$kb" >> $O/tune_euler.log
BLF_CCM_TUNE_FBD_TILE_KB=$kb DYN_NC=none DYN_EULER_ONLY=1 timeout 300 python tools/tune.py --dyn-only >> $O/tune_euler.log 2>&1
done
echo "== BLF_CCM_TUNE_FBD_NO_BULK=1 (per-thread copies)" >> $O/tune_euler.log
BLF_CCM_TUNE_FBD_NO_BULK=1 DYN_NC=none DYN_EULER_ONLY=1 timeout 300 python tools/tune.py --dyn-only >> $O/tune_euler.log 2>&1
grep -E "==|Euler" $O/tune_euler.log | cut -c1-200
DYN_NC=none DYN_EULER_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sys_fbd_euler -s 6 -c 1 -f -o $O/prof_fbd_euler python tools/tune.py --dyn-only > $O/ncu_fbd_euler.log 2>&1
ncu -i $O/prof_fbd_euler.ncu-rep --page details > $O/prof_fbd_euler.details.txt 2>/dev/null
ncu -i $O/prof_fbd_euler.ncu-rep --page raw --csv > $O/prof_fbd_euler.raw.csv 2>/dev/null
rm -f $O/prof_fbd_euler.ncu-rep
grep -E "Duration|DRAM Throughput|Registers Per|Achieved Occupancy|Theoretical Occupancy|Issue Slots Busy|No Eligible|Executed Ipc|Memory Throughput|Block Limit" $O/prof_fbd_euler.details.txt | head -20
ls -la $O

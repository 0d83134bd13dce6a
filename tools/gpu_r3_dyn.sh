#!/bin/bash
# Round 2, third session: the mass-matrix solve / whole dynamics step -- GPU tests, sweep, ncu of the nc = 29 kernel.
TAG=${1:-r3dyn}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dyn.py -m gpu -q -x -rs --durations=5 > $O/pytest_dyn.log 2>&1; echo "pytest exit $?" >> $O/pytest_dyn.log
grep -E "^E  .*|passed|failed|^FAILED|SKIPPED|pytest exit" $O/pytest_dyn.log | cut -c1-220 | tail -12
timeout 600 python tools/tune.py --dyn-only > $O/tune_dyn.log 2>&1; echo "tune exit $?" >> $O/tune_dyn.log
cat $O/tune_dyn.log | cut -c1-230
if [ "${NCU:-1}" = "1" ]; then
DYN_NC=29 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ccm_llt_solve -s 6 -c 1 -f -o $O/prof_llt29 python tools/tune.py --dyn-only > $O/ncu_llt29.log 2>&1
ncu -i $O/prof_llt29.ncu-rep --page details > $O/prof_llt29.details.txt 2>/dev/null
ncu -i $O/prof_llt29.ncu-rep --page raw --csv > $O/prof_llt29.raw.csv 2>/dev/null
rm -f $O/prof_llt29.ncu-rep
grep -E "Duration|DRAM Throughput|Registers Per|Achieved Occupancy|Theoretical Occupancy|Issue Slots Busy|No Eligible|Shared Memory Configuration|L1/TEX Hit|Executed Ipc|Block Limit" $O/prof_llt29.details.txt | head -30
fi
ls -la $O

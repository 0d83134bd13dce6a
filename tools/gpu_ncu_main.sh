#!/bin/bash
# ncu evidence for the headline kernel: launch list + one full capture, exported to CSV on the box.
TAG=${1:-ncumain}
O=gpurun_out/$TAG
mkdir -p $O
PROF="python bench.py --steps 20 --warmup 3 --only-main"
$PROF > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv $PROF > $O/ncu_launches.log 2>&1
$PROF > $O/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ccm_soa_kernel -s 5 -c 1 -f -o $O/prof_soa $PROF > $O/ncu_full.log 2>&1
ncu -i $O/prof_soa.ncu-rep --page raw --csv > $O/prof_soa.raw.csv 2>/dev/null
ncu -i $O/prof_soa.ncu-rep --page details > $O/prof_soa.details.txt 2>/dev/null
rm -f $O/prof_soa.ncu-rep
tail -2 $O/ncu_full.log; ls -la $O

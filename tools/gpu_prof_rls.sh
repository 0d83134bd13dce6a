#!/bin/bash
TAG=${1:-prls}; O=gpurun_out/$TAG; mkdir -p $O
for v in 0 1; do
  BLF_CCM_TUNE_RLS_PIPE=$v python tools/prof_rls.py > $O/plain_$v.log 2>&1 && \
  BLF_CCM_TUNE_RLS_PIPE=$v ncu --set full --clock-control none --import-source on -k regex:rls_advance -s 5 -c 1 -o $O/rls_$v python tools/prof_rls.py > $O/ncu_$v.log 2>&1
  ncu -i $O/rls_$v.ncu-rep --page details > $O/rls_$v.details.txt 2>&1
  ncu -i $O/rls_$v.ncu-rep --page raw --csv > $O/rls_$v.raw.csv 2>&1
  rm -f $O/rls_$v.ncu-rep
done
ls -la $O

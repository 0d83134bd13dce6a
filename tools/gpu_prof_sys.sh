#!/bin/bash
# ncu captures of the System-row kernels: bash tools/gpu_prof_sys.sh <tag> [workloads...]
# The .ncu-rep files are exported to CSV on the box and removed (gpurun_out/ is capped at 64 MiB).
TAG=${1:-profsys}; shift
WL=${@:-rollout_small rollout_large rollout_large_rho rollout_full genforce genforce_small}
O=gpurun_out/$TAG
mkdir -p $O
for W in $WL; do
  case $W in rollout_small) K=ccm_rollout;; rollout*) K=ccm_rollout_kernel;; genforce_6x2|genforce_12x2|genforce_narrow) K=ccm_genforce_packed_kernel;; genforce*) K=ccm_genforce_kernel;; *) K=sys_kin_euler_kernel;; esac
  python tools/prof_sys.py $W > $O/plain_$W.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $O/$W python tools/prof_sys.py $W > $O/ncu_$W.log 2>&1
  tail -1 $O/ncu_$W.log
  ncu -i $O/$W.ncu-rep --page raw --csv > $O/$W.raw.csv 2>/dev/null
  ncu -i $O/$W.ncu-rep --page details > $O/$W.details.txt 2>/dev/null
  ncu -i $O/$W.ncu-rep --page source --csv > $O/$W.source.csv 2>/dev/null
  rm -f $O/$W.ncu-rep
done
ls -la $O

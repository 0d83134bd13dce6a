#!/bin/bash
# configs[1] evidence: bench line, then (after it exited 0) ncu launch list + one full capture.
TAG=${1:-c2}
O=gpurun_out/$TAG
mkdir -p $O
python bench.py --workload config2 > $O/bench_config2.json 2> $O/bench_config2.err; echo "exit $?" >> $O/bench_config2.err
cut -c1-1800 $O/bench_config2.json; tail -2 $O/bench_config2.err
PROF="python bench.py --workload config2 --steps 30 --warmup 3"
$PROF > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ccm_ -c 40 --csv --log-file $O/launches_config2.csv $PROF > $O/ncu_launches.log 2>&1
$PROF > $O/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ccm_soa_kernel -s 5 -c 1 -o $O/prof_soa_wrench $PROF > $O/ncu_full.log 2>&1
ncu -i $O/prof_soa_wrench.ncu-rep --page details > $O/soa_wrench.details.txt 2>&1
ncu -i $O/prof_soa_wrench.ncu-rep --page raw --csv > $O/soa_wrench.raw.csv 2>&1
tail -3 $O/ncu_full.log; ls -la $O

#!/bin/bash
# One gpurun call: parity tests, smoke, bench, variant sweep, ncu launch list + full capture.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh <tag>
TAG=${1:-r1}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv > $O/gpu.csv 2>&1
nproc > $O/nproc.txt; lscpu | head -20 >> $O/nproc.txt
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
tail -2 $O/smoke.log
python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?" >> $O/bench.err
cat $O/bench.json | cut -c1-1500
timeout 900 python tools/tune.py > $O/tune.log 2>&1; echo "tune exit $?" >> $O/tune.log
cat $O/tune.log | tail -60
# ncu: only after the identical command exited 0 without ncu
PROF="python bench.py --steps 20 --warmup 3 --only-main"
$PROF > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv $PROF > $O/ncu_launches.log 2>&1
$PROF > $O/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ccm_soa_kernel -s 5 -c 2 -o $O/prof_soa $PROF > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log
ls -la $O

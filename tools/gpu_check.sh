#!/bin/bash
# Full GPU check: all -m gpu tests, smoke, both bench arms.  bash tools/gpu_check.sh <tag>
TAG=${1:-check}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv > $O/gpu.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E "^E  .*(assert|Error)|passed|failed|^FAILED|pytest exit" $O/pytest_gpu.log | cut -c1-220 | tail -15
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
tail -2 $O/smoke.log
python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?" >> $O/bench.err
tail -3 $O/bench.err; cat $O/bench.json | cut -c1-4000

#!/bin/bash
# Round 2, first GPU call: all GPU tests, smoke, both bench arms, the low-memory fallback of the
# headline, the facade latency.   bash tools/gpu_r2a.sh <tag>
TAG=${1:-r2a}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv > $O/gpu.csv 2>&1
nproc > $O/nproc.txt; lscpu | head -25 >> $O/nproc.txt; free -g >> $O/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E "^E  .*(assert|Error)|passed|failed|^FAILED|pytest exit|^[0-9.]+s (call|setup)" $O/pytest_gpu.log | cut -c1-220 | tail -30
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
tail -2 $O/smoke.log
python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2> $O/bench.time; echo "bench exit $?" >> $O/bench.err
tail -3 $O/bench.err; cat $O/bench.time; cut -c1-900 $O/bench.json
python bench.py --steps 3 --warmup 3 --only-main --force-windows 2 > $O/bench_windows2.json 2> $O/bench_windows2.err; echo "windows exit $?" >> $O/bench_windows2.err
cut -c1-400 $O/bench_windows2.json
./bipedal_locomotion_framework_b200/lib/ContinuousContactModelUnitTests > $O/cpp_ccm_tests.log 2>&1; echo "cpp exit $?" >> $O/cpp_ccm_tests.log
grep -E "per-instance|exit|passed|failed" $O/cpp_ccm_tests.log | tail -5
ls -la $O

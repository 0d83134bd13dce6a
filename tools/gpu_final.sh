#!/bin/bash
# Final evidence of a session: all GPU tests, smoke, both bench arms, configs[1] line + launch list,
# variant sweep.  bash tools/gpu_final.sh <tag>
TAG=${1:-final}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv > $O/gpu.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E "^E  .*(assert|Error)|passed|failed|^FAILED|pytest exit" $O/pytest_gpu.log | cut -c1-220 | tail -15
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
tail -2 $O/smoke.log
python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?" >> $O/bench.err
tail -1 $O/bench.err; cut -c1-600 $O/bench.json
python bench.py --workload config2 > $O/bench_config2.json 2> $O/bench_config2.err; echo "exit $?" >> $O/bench_config2.err
cut -c1-300 $O/bench_config2.json
PROF="python bench.py --workload config2 --steps 30 --warmup 3"
$PROF > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ccm_ -c 40 --csv --log-file $O/launches_config2.csv $PROF > $O/ncu_launches.log 2>&1
timeout 300 python tools/tune.py > $O/tune.log 2>&1; echo "tune exit $?" >> $O/tune.log
tail -45 $O/tune.log

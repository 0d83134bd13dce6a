#!/bin/bash
# Large single-GPU workloads: bash tools/gpu_large.sh <tag>
TAG=${1:-large}
O=gpurun_out/$TAG
mkdir -p $O
python -m pytest tests -m gpu -q -k "config4_full_size" > $O/pytest_config4_full.log 2>&1; tail -3 $O/pytest_config4_full.log
python bench.py --workload config4 --steps 20 --warmup 3 > $O/bench_config4_n1.json 2> $O/bench_config4_n1.err; echo "config4 exit $?"; tail -c 900 $O/bench_config4_n1.json; tail -3 $O/bench_config4_n1.err
python bench.py --workload config5 --steps 10 --warmup 3 > $O/bench_config5_n1.json 2> $O/bench_config5_n1.err; echo "config5 exit $?"; tail -c 900 $O/bench_config5_n1.json; tail -3 $O/bench_config5_n1.err

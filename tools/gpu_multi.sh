#!/bin/bash
# Multi-GPU bench on one box: bash tools/gpu_multi.sh <N> <tag>
N=${1:-2}; TAG=${2:-multi}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $O/gpus.csv 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
for n in $(echo $N | tr ',' ' '); do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 1000 --warmup 10 > $O/bench_n1.json 2> $O/bench_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 1000 --warmup 10 > $O/bench_n$n.json 2> $O/bench_n$n.err
  fi
  echo "N=$n exit $?"; tail -c 1200 $O/bench_n$n.json; echo; tail -3 $O/bench_n$n.err
done

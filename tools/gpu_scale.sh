#!/bin/bash
# Scaling sweep on one 8-GPU box: bash tools/gpu_scale.sh <tag>
TAG=${1:-scale}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $O/gpus.csv 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
run() {  # run <N> <workload> <steps>
  local n=$1 wl=$2 steps=$3 out=$O/bench_${2}_n$1
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --workload $wl --steps $steps --warmup 5 --no-cpu > $out.json 2> $out.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --workload $wl --steps $steps --warmup 5 --no-cpu > $out.json 2> $out.err
  fi
  echo "$wl N=$n exit $? $(python - <<PY
import json
try:
    d=[json.loads(l) for l in open('$out.json') if l.startswith('{')][-1]
    print('value %.3f G evals/s, %.4f ms/step' % (d['value']/1e9, d['ms_per_step']), 'mpc %.3f G' % (d['mpc']['value']/1e9) if d.get('mpc') else '', d.get('parity'))
except Exception as e:
    print('no json', e)
PY
)"
}
for n in 1 2 4 8; do run $n config3 1000; done
for n in 2 4 8; do run $n config4 20; done
for n in 2 4 8; do run $n config5 10; done
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 8 --workload config4 --steps 5 --warmup 3 --no-cpu 2>&1 | grep -iE "NVLS|NET/|Channel 00.*via|Connected all" | head -12 > $O/nccl_info.txt

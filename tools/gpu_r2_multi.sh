#!/bin/bash
# Round 2 multi-GPU call: bash tools/gpu_r2_multi.sh <tag> "<list of N>"   (under gpurun --gpus max(N))
TAG=${1:-r2m}
NS=${2:-"2"}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $O/gpus.csv 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
nproc > $O/nproc.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> $O/nproc.txt
timeout 900 python -m pytest tests/test_gpu_p2p.py tests/test_gpu_parity.py tests/test_gpu_sys.py -m gpu -q -x -k "p2p or restore or warp_specialised" > $O/pytest_p2p.log 2>&1; echo "pytest exit $?" >> $O/pytest_p2p.log
tail -4 $O/pytest_p2p.log
for n in $NS; do
  out=$O/bench_n$n
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 20 --warmup 5 > $out.json 2> $out.err ) 2> $out.time
  echo "N=$n exit $? $(grep real $out.time)"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('$out.json') if l.startswith('{')][-1]
    print('  value %.3f G evals/s, %.4f ms/step, frac %.3f, sustained %s' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('sustained_frac')))
    print('  parity', d['parity']['ok'], 'exchange_check', d['exchange_check'] and {k: d['exchange_check'][k] for k in ('steps','ok','all_ok','distinct_costs','blocked_peer_probe')})
    h=d['configs3_het64m']; print('  het64m %.3f G frac %.3f parity %s xchg %s' % (h['value']/1e9, h['roofline']['frac'], h['parity']['ok'], h['exchange_check'] and h['exchange_check']['all_ok']))
    m=d['configs2']['mpc']; f=d['configs2']['mpc_fused']; print('  mpc %.1f us (xchg %s)  fused %.1f us  eval_only %.3f G' % (m['ms_per_step']*1e3, m['exchange_check'] and m['exchange_check']['all_ok'], f['ms_per_step']*1e3, d['configs2']['eval_only']['value']/1e9))
    e=d['e2e']; print('  e2e %.1f M evals/s, ceiling frac %.2f, dense %.1f M' % (e['value']/1e6, e['pcie']['frac_of_ceiling'], e['dense_download']['value']/1e6))
except Exception as ex:
    print('  no json', ex)
PY
  tail -3 $out.err
done
python bench.py --impl reference --gpus 1 --steps 10 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
ls -la $O

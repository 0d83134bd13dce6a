#!/bin/bash
# Round 2, fourth session: bench.py at N = 8 with the final sources (one call, bench only).
TAG=${1:-r4n8}
O=gpurun_out/$TAG
mkdir -p $O
n=${2:-8}
out=$O/bench_n$n
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 20 --warmup 5 > $out.json 2> $out.err ) 2> $out.time
echo "N=$n exit $? $(grep real $out.time)"
python - <<PY
import json
d=[json.loads(l) for l in open('$out.json') if l.startswith('{')][-1]
print('  value %.3f G evals/s, %.4f ms/step, frac %.3f, sustained %s' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('sustained_frac')))
print('  parity', d['parity']['ok'], 'exchange_check', d['exchange_check'] and {k: d['exchange_check'][k] for k in ('steps','ok','all_ok','distinct_costs','blocked_peer_probe')})
h=d['configs3_het64m']; print('  het64m %.3f G frac %.3f parity %s xchg %s' % (h['value']/1e9, h['roofline']['frac'], h['parity']['ok'], h['exchange_check'] and h['exchange_check']['all_ok']))
m=d['configs2']['mpc']; f=d['configs2']['mpc_fused']; print('  mpc %.1f us (xchg %s)  fused %.1f us' % (m['ms_per_step']*1e3, m['exchange_check'] and m['exchange_check']['all_ok'], f['ms_per_step']*1e3))
e=d['e2e']; print('  e2e %.1f M evals/s, ceiling frac %.2f' % (e['value']/1e6, e['pcie']['frac_of_ceiling']))
PY
tail -2 $out.err

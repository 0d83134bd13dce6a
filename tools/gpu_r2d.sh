#!/bin/bash
# Round 2, fourth GPU call: all GPU tests (new rollout kernel, general RLS, packed J^T wrench), rollout
# sweep, bench.
TAG=${1:-r2d}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x --durations=10 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E "^E  .*|passed|failed|^FAILED|pytest exit|^[0-9.]+s (call|setup)" $O/pytest_gpu.log | cut -c1-220 | tail -30
timeout 900 python tools/rollout_sweep.py > $O/rollout_sweep.log 2>&1; cat $O/rollout_sweep.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -2 $O/smoke.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2> $O/bench.time; echo "bench exit $?" >> $O/bench.err
tail -3 $O/bench.err; cat $O/bench.time
TAG=$TAG python - <<'PY'
import json
d=json.loads(open("gpurun_out/%s/bench.json" % __import__("os").environ["TAG"]).read())
print("value %.3f G, frac %.3f, sustained %.3f" % (d["value"]/1e9, d["roofline"]["frac"], d["roofline"]["sustained_frac"]))
print("mpc %.1f us, mpc_fused %.1f us, e2e %.1f M (dense %.1f M, ceiling frac %.2f)" % (d["configs2"]["mpc"]["ms_per_step"]*1e3, d["configs2"]["mpc_fused"]["ms_per_step"]*1e3, d["e2e"]["value"]/1e6, d["e2e"]["dense_download"]["value"]/1e6, d["e2e"]["pcie"]["frac_of_ceiling"]))
print({k: round(v["hbm_frac_of_measured"],3) for k,v in d["next_rows"].items()})
PY
ls -la $O

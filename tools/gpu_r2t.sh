#!/bin/bash
# Round 2 final single-GPU validation: all GPU tests (with skip reasons), smoke, both bench arms, C++ tests,
# ncu of the final fifth rollout form.
TAG=${1:-r2t}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit,memory.total --format=csv > $O/gpu.csv 2>&1
nproc > $O/nproc.txt
timeout 1800 python -m pytest tests -m gpu -q -x -rs --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E "^E  .*|passed|failed|^FAILED|SKIPPED|pytest exit" $O/pytest_gpu.log | cut -c1-220 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -2 $O/smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2> $O/bench.time; echo "bench exit $?" >> $O/bench.err
tail -3 $O/bench.err; cat $O/bench.time
TAG=$TAG python - <<'PY'
import json, os
d=json.loads(open("gpurun_out/%s/bench.json" % os.environ["TAG"]).read())
r=json.loads(open("gpurun_out/%s/bench_reference.json" % os.environ["TAG"]).read())
print("value %.3f G, frac %.3f, sustained %.3f, traffic %s" % (d["value"]/1e9, d["roofline"]["frac"], d["roofline"]["sustained_frac"], d["roofline"]["traffic"]))
print("mpc %.1f us, mpc_fused %.1f us, e2e %.1f M (dense %.1f M, ceiling frac %.2f), cpu %.1f M (1 thread %.2f M), reference arm %.1f M" % (d["configs2"]["mpc"]["ms_per_step"]*1e3, d["configs2"]["mpc_fused"]["ms_per_step"]*1e3, d["e2e"]["value"]/1e6, d["e2e"]["dense_download"]["value"]/1e6, d["e2e"]["pcie"]["frac_of_ceiling"], d["cpu_baseline"]["value"]/1e6, d["cpu_baseline"]["single_thread_value"]/1e6, r["value"]/1e6))
print({k: round(v["hbm_frac_of_measured"],3) for k,v in d["next_rows"].items()})
PY
./bipedal_locomotion_framework_b200/lib/ContinuousContactModelUnitTests > $O/cpp_ccm_tests.log 2>&1; tail -2 $O/cpp_ccm_tests.log
python tools/prof_rollout.py 0.01 > $O/prof_rollout_plain.log 2>&1 && cat $O/prof_rollout_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:ccm_rollout_ws5 -s 4 -c 1 -f -o $O/prof_ws5 python tools/prof_rollout.py 0.01 > $O/ncu_ws5.log 2>&1
ncu -i $O/prof_ws5.ncu-rep --page details > $O/prof_ws5.details.txt 2>/dev/null
rm -f $O/prof_ws5.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ccm_ -c 30 --csv --log-file $O/launches_rollout.csv python tools/prof_rollout.py 0.01 > $O/ncu_l.log 2>&1
grep ws5 $O/launches_rollout.csv | tail -2 | cut -c1-250
ls -la $O

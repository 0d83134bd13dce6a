#!/bin/bash
# Round 2, fourth session: full single-GPU validation -- all GPU tests (with skip reasons), smoke, both bench arms, C++ tests.
TAG=${1:-r4full}; export TAG
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit,memory.total --format=csv > $O/gpu.csv 2>&1
nproc > $O/nproc.txt
timeout 2400 python -m pytest tests -m gpu -q -x -rs --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E "^E  .*|passed|failed|^FAILED|SKIPPED|pytest exit" $O/pytest_gpu.log | cut -c1-220 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -2 $O/smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2> $O/bench.time; echo "bench exit $?" >> $O/bench.err
tail -3 $O/bench.err; cat $O/bench.time
TAG=$TAG python - <<'PY'
import json, os
d=json.loads(open("gpurun_out/%s/bench.json" % os.environ["TAG"]).read())
r=json.loads(open("gpurun_out/%s/bench_reference.json" % os.environ["TAG"]).read())
print("value %.3f G, frac %.3f, sustained %.3f" % (d["value"]/1e9, d["roofline"]["frac"], d["roofline"]["sustained_frac"]))
print("mpc %.1f us, mpc_fused %.1f us, e2e %.1f M (ceiling frac %.2f), cpu %.1f M, reference arm %.1f M" % (d["configs2"]["mpc"]["ms_per_step"]*1e3, d["configs2"]["mpc_fused"]["ms_per_step"]*1e3, d["e2e"]["value"]/1e6, d["e2e"]["pcie"]["frac_of_ceiling"], d["cpu_baseline"]["value"]/1e6, r["value"]/1e6))
print({k: round(v["hbm_frac_of_measured"],3) for k,v in d["next_rows"].items()})
for k in ("floating_base_dynamics_29", "floating_base_dynamics_12"):
    v = d["next_rows"][k]
    print(k, "solve %.1f M/s, whole step %.1f M/s, Euler step %.1f M/s (%.3f of HBM), cpu port %.2f M/s on %d cores" % (v["solve_systems_per_s"]/1e6, v["whole_step_systems_per_s"]/1e6, v["euler_step_systems_per_s"]/1e6, v["euler_step_hbm_frac"], v["cpu_baseline"]["value"]/1e6, v["cpu_baseline"]["cores"]))
PY
for t in ContinuousContactModelUnitTests IntegratorUnitTests RecursiveLeastSquareUnitTests FloatingBaseSystemDynamicsUnitTests; do ./bipedal_locomotion_framework_b200/lib/$t > $O/cpp_$t.log 2>&1; echo "$t exit $?"; tail -1 $O/cpp_$t.log; done
ls $O
# launch list of the dynamics rows (solve, J^T wrench, Euler step) under ncu: per-launch durations, cold and serialised
DYN_NC=29 timeout 300 python tools/tune.py --dyn-only > $O/tune_dyn29.log 2>&1 && \
DYN_NC=29 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_dyn29.csv python tools/tune.py --dyn-only > $O/ncu_dyn29.log 2>&1
grep -E "acceleration|Euler|solve" $O/tune_dyn29.log | cut -c1-170
python - <<'PY'
import csv, os, collections
O = "gpurun_out/%s" % os.environ.get("TAG", "r4full")
rows = [r for r in csv.reader(open(O + "/launches_dyn29.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4].split("(")[0][:70]].append(float(r[-1]))
for k, v in agg.items():
    print("%-72s launches %4d  mean %10.1f us" % (k, len(v), sum(v) / len(v) / 1e3))
PY

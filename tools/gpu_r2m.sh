#!/bin/bash
# Round 2 full check on one GPU: all GPU tests, smoke, bench (default command), ncu launch list of the
# bench's headline loop, ncu --set full of the headline kernel (dram bytes -> roofline.traffic) and of
# the fifth rollout form.
TAG=${1:-r2m}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.csv 2>&1
timeout 1800 python -m pytest tests -m gpu -q -x --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E "^E  .*|passed|failed|^FAILED|pytest exit" $O/pytest_gpu.log | cut -c1-220 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -2 $O/smoke.log
python bench.py --impl reference --gpus 1 --steps 10 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2> $O/bench.time; echo "bench exit $?" >> $O/bench.err
tail -3 $O/bench.err; cat $O/bench.time
TAG=$TAG python - <<'PY'
import json, os
d=json.loads(open("gpurun_out/%s/bench.json" % os.environ["TAG"]).read())
print("value %.3f G, frac %.3f, sustained %.3f, traffic %s" % (d["value"]/1e9, d["roofline"]["frac"], d["roofline"]["sustained_frac"], d["roofline"]["traffic"]))
print("mpc %.1f us, mpc_fused %.1f us, e2e %.1f M (dense %.1f M, ceiling frac %.2f), cpu %.1f M" % (d["configs2"]["mpc"]["ms_per_step"]*1e3, d["configs2"]["mpc_fused"]["ms_per_step"]*1e3, d["e2e"]["value"]/1e6, d["e2e"]["dense_download"]["value"]/1e6, d["e2e"]["pcie"]["frac_of_ceiling"], d["cpu_baseline"]["value"]/1e6))
print({k: round(v["hbm_frac_of_measured"],3) for k,v in d["next_rows"].items()})
PY
# launch list of the bench's own headline loop (same command, headline only)
PROF="python bench.py --gpus 1 --steps 20 --warmup 5 --only-main"
$PROF > $O/plain_main.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_main.csv $PROF > $O/ncu_launches.log 2>&1
grep -c ccm_ $O/launches_main.csv
# full capture of the headline kernels at 2^23 states
python tools/prof_headline.py 23 > $O/prof_headline_plain.log 2>&1 && cat $O/prof_headline_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:ccm_soa_kernel -s 3 -c 1 -f -o $O/prof_soa_cost python tools/prof_headline.py 23 > $O/ncu_soa_cost.log 2>&1
ncu -i $O/prof_soa_cost.ncu-rep --page raw --csv > $O/prof_soa_cost.raw.csv 2>/dev/null
ncu -i $O/prof_soa_cost.ncu-rep --page details > $O/prof_soa_cost.details.txt 2>/dev/null
python tools/prof_headline.py 23 het > $O/prof_headline_het_plain.log 2>&1 && cat $O/prof_headline_het_plain.log && \
ncu --set full --clock-control none -k regex:ccm_soa_kernel -s 3 -c 1 -f -o $O/prof_soa_cost_het python tools/prof_headline.py 23 het > $O/ncu_soa_cost_het.log 2>&1
ncu -i $O/prof_soa_cost_het.ncu-rep --page raw --csv > $O/prof_soa_cost_het.raw.csv 2>/dev/null
rm -f $O/prof_soa_cost_het.ncu-rep
# the fifth rollout form (automatic choice at the MPC size)
python tools/prof_rollout.py 0.01 > $O/prof_rollout_plain.log 2>&1 && cat $O/prof_rollout_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:ccm_rollout_ws5 -s 4 -c 1 -f -o $O/prof_ws5 python tools/prof_rollout.py 0.01 > $O/ncu_ws5.log 2>&1
ncu -i $O/prof_ws5.ncu-rep --page details > $O/prof_ws5.details.txt 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file $O/launches_rollout.csv python tools/prof_rollout.py 0.01 > $O/ncu_l.log 2>&1
SWEEP_ONLY=auto,ws3,ws5 timeout 600 python tools/rollout_sweep.py > $O/rollout_sweep.log 2>&1; grep -E "===|auto" $O/rollout_sweep.log
ls -la $O

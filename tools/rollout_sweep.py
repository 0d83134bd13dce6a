#!/usr/bin/env python
"""Fused-rollout kernel variants at the sampling-MPC size (2 feet x 4096 samples x 100 steps = 8 192
chains) and two larger sizes: per-step time (rollout kernel + reduction launch) for every variant
BLF_CCM_TUNE_ROLLOUT_WS / _SPLIT can force, with and without the Baumgarte term; all variants must
give the same arg-min and costs within 1e-12.   python tools/rollout_sweep.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from bipedal_locomotion_framework_b200 import synthetic as syn
from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch
from bipedal_locomotion_framework_b200.system import RolloutBatch

FEET, H = 2, 100
REF, WTS = [0.0, 0.0, 30.0, 0.0, 0.0, 0.0], [1.0, 10.0]


def timeit(fn, iters=200, warm=20):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


VARIANTS = [("auto", {}), ("split 1", {"BLF_CCM_TUNE_ROLLOUT_SPLIT": 1, "BLF_CCM_TUNE_ROLLOUT_WS": 2}),
            ("split 2", {"BLF_CCM_TUNE_ROLLOUT_SPLIT": 2, "BLF_CCM_TUNE_ROLLOUT_WS": 2}),
            ("split 4", {"BLF_CCM_TUNE_ROLLOUT_SPLIT": 4, "BLF_CCM_TUNE_ROLLOUT_WS": 2}),
            ("ws2 C=3 (cp.async ring, 2 launches)", {"BLF_CCM_TUNE_ROLLOUT_WS": 3}),
            ("ws2 C=7", {"BLF_CCM_TUNE_ROLLOUT_WS": 7}),
            ("ws3 C=3 (TMA twists, fused reduction)", {"BLF_CCM_TUNE_ROLLOUT_WS": 13}),
            ("ws3 C=7", {"BLF_CCM_TUNE_ROLLOUT_WS": 17}),
            ("ws4 C=3 (4 lanes per chain, TMA, fused reduction)", {"BLF_CCM_TUNE_ROLLOUT_WS": 23}),
            ("ws4 C=5", {"BLF_CCM_TUNE_ROLLOUT_WS": 25}), ("ws4 C=7", {"BLF_CCM_TUNE_ROLLOUT_WS": 27}),
            ("ws5 layout 0 (box hand-over, loader warp, 3 consumers)", {"BLF_CCM_TUNE_ROLLOUT_WS": 33}),
            ("ws5 layout 1 (4 consumers, steps 3/3/1/1)", {"BLF_CCM_TUNE_ROLLOUT_WS": 34}), ("ws5 layout 2 (4 consumers, steps 2/2/2/2)", {"BLF_CCM_TUNE_ROLLOUT_WS": 35})]
if os.environ.get("SWEEP_ONLY"):   # e.g. SWEEP_ONLY=auto,ws3,ws5
    keep = os.environ["SWEEP_ONLY"].split(",")
    VARIANTS = [v for v in VARIANTS if any(v[0].startswith(k) for k in keep)]

for samples in (4096, 1024, 16384, 65536):
    chains = FEET * samples
    n = chains * H
    st = syn.make_states(min(n, 1 << 20), seed=45)
    reps = (n + st["n"] - 1) // st["n"]
    pl = np.tile(syn.aos_to_planes(st["twists"], st["poses"], st["null_poses"]), (1, reps))[:, :n]
    for rho in (0.0, 0.01):
        print(f"=== {samples} samples x {FEET} feet x {H} steps = {n} evaluations, rho = {rho} ===", flush=True)
        base_cost = None
        for name, env in VARIANTS:
            for k in ("BLF_CCM_TUNE_ROLLOUT_SPLIT", "BLF_CCM_TUNE_ROLLOUT_WS"):
                os.environ.pop(k, None)
            for k, v in env.items():
                os.environ[k] = str(v)
            b = ContinuousContactModelBatch(0)
            b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
            rb = RolloutBatch(b)
            sets = [torch.from_numpy(np.ascontiguousarray(pl)).cuda() for _ in range(3)]
            calls = [rb.prepare(samples, FEET, H, 0.01, rho, p[0:6], p[6:9, :chains], p[9:18, :chains],
                                p[18:30, :chains], REF, WTS, mask=0, want_cost=True) for p in sets]
            ms = timeit(lambda i: calls[i % 3][0]())
            cost = calls[0][1]["cost"].cpu().numpy()
            best = b.decode_best(calls[0][1]["best"])
            if base_cost is None:
                base_cost = cost
            err = float(np.max(np.abs(cost - base_cost) / np.maximum(np.abs(base_cost), 1e-300)))
            print(f"  {name:52s} {ms*1e3:8.1f} us/step  {n/ms/1e6:8.2f} G evals/s  argmin {best[1]:6d}  "
                  f"max cost dev vs first variant {err:.1e}", flush=True)
            del b, rb, calls, sets

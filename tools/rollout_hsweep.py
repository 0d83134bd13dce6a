#!/usr/bin/env python
"""Fixed cost vs per-step cost of the fused rollout at small batches: time one variant over a range
of horizons and fit  t(H) = t0 + H * dt.   BLF_CCM_TUNE_ROLLOUT_WS=13 python tools/rollout_hsweep.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from bipedal_locomotion_framework_b200 import synthetic as syn
from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch
from bipedal_locomotion_framework_b200.system import RolloutBatch

FEET = 2
REF, WTS = [0.0, 0.0, 30.0, 0.0, 0.0, 0.0], [1.0, 10.0]
HS = (1, 8, 24, 48, 100, 200, 400)


def timeit(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


variants = [v for v in os.environ.get("HSWEEP_VARIANTS", "13").split(",")]
for var in variants:
    os.environ["BLF_CCM_TUNE_ROLLOUT_WS"] = var
    for samples in (1024, 4096):
        chains = FEET * samples
        for rho in (0.0, 0.01):
            ts = []
            for H in HS:
                n = chains * H
                st = syn.make_states(min(n, 1 << 18), seed=45)
                reps = (n + st["n"] - 1) // st["n"]
                pl = np.tile(syn.aos_to_planes(st["twists"], st["poses"], st["null_poses"]), (1, reps))[:, :n]
                b = ContinuousContactModelBatch(0)
                b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
                rb = RolloutBatch(b)
                p = torch.from_numpy(np.ascontiguousarray(pl)).cuda()
                call, out = rb.prepare(samples, FEET, H, 0.01, rho, p[0:6], p[6:9, :chains], p[9:18, :chains],
                                       p[18:30, :chains], REF, WTS, mask=0, want_cost=True)
                ts.append(timeit(call))
                del b, rb, call, out, p
            A = np.vstack([np.ones(len(HS)), np.array(HS, float)]).T
            t0, dt = np.linalg.lstsq(A[2:], np.array(ts[2:]), rcond=None)[0]
            print(f"variant {var} samples {samples} rho {rho}: " + " ".join(f"H={h}:{t:.1f}" for h, t in zip(HS, ts)) +
                  f" us | fit (H>=24) t0 {t0:.1f} us + {dt*1e3:.0f} ns/step", flush=True)

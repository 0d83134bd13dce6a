#!/usr/bin/env python
"""A few fused-rollout steps of ONE kernel variant at the sampling-MPC size, for ncu:
   BLF_CCM_TUNE_ROLLOUT_WS=13 python tools/prof_rollout.py [rho] [samples]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from bipedal_locomotion_framework_b200 import synthetic as syn
from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch
from bipedal_locomotion_framework_b200.system import RolloutBatch

rho = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
FEET, H = 2, 100
chains = FEET * samples
n = chains * H
st = syn.make_states(min(n, 1 << 20), seed=45)
reps = (n + st["n"] - 1) // st["n"]
pl = np.tile(syn.aos_to_planes(st["twists"], st["poses"], st["null_poses"]), (1, reps))[:, :n]
b = ContinuousContactModelBatch(0)
b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
rb = RolloutBatch(b)
p = torch.from_numpy(np.ascontiguousarray(pl)).cuda()
call, out = rb.prepare(samples, FEET, H, 0.01, rho, p[0:6], p[6:9, :chains], p[9:18, :chains], p[18:30, :chains],
                       [0.0, 0.0, 30.0, 0.0, 0.0, 0.0], [1.0, 10.0], mask=0, want_cost=True)
for _ in range(6):
    call()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    call()
e.record()
torch.cuda.synchronize()
print(f"rho {rho} samples {samples} variant {os.environ.get('BLF_CCM_TUNE_ROLLOUT_WS', 'auto')}: "
      f"{a.elapsed_time(e) / 50 * 1e3:.1f} us/step, argmin {b.decode_best(out['best'])}")

#!/usr/bin/env python
"""The headline step (blf_ccm_rollout_cost_argmin_soa: ccm_soa_kernel<WRENCH|AUTODYN|CTRL, uniform,
COST> + ccm_cost_reduce_kernel) at a size ncu can replay: 2^23 states = 5 GB of planes, 3.2 GB of
outputs (at the bench's 2^28 states ncu would have to save and restore 103 GB per pass).
   python tools/prof_headline.py [log2 n] [het]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from bipedal_locomotion_framework_b200 import synthetic as syn
from bipedal_locomotion_framework_b200.contact_models import FULL, ContinuousContactModelBatch

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 23
het = len(sys.argv) > 2 and sys.argv[2] == "het"
n, L = 1 << lg, 256
b = ContinuousContactModelBatch(0)
b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
planes, prm = syn.make_planes_torch(n, torch.device("cuda:0"), seed=46, heterogeneous=het)
call, out, cost, best = b.prepare_rollout(planes, L, [0.0, 0.0, 30.0, 0.0, 0.0, 0.0], [1.0, 10.0], param_planes=prm,
                                          mask=FULL)
for _ in range(3):
    call()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    call()
e.record()
torch.cuda.synchronize()
ms = a.elapsed_time(e) / 10
bpe = 632 if het else 600
print(f"n = 2^{lg}{' het' if het else ''}: {ms*1e3:.1f} us/step, {n/ms/1e6:.2f} G evals/s, {bpe*n/ms/1e6:.0f} GB/s algorithmic, "
      f"argmin {b.decode_best(best)}")

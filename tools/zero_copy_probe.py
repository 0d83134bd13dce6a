#!/usr/bin/env python
"""Can the AoS kernel work straight on pinned host memory (UVA: the pinned pointer is a device pointer
too), i.e. SM-issued PCIe reads/writes instead of copy-engine transfers?  Times blf_ccm_eval_batch_aos
with (a) all arrays in pinned host memory, (b) inputs in host memory / outputs in HBM, (c) inputs in
HBM / outputs in host memory, against blf_ccm_eval_batch_host on the same buffers.
   python tools/zero_copy_probe.py [n]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from bipedal_locomotion_framework_b200 import synthetic as syn
from bipedal_locomotion_framework_b200.contact_models import FULL, WRENCH, AUTODYN, ContinuousContactModelBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
st = syn.make_states(n, seed=46)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
tw, po, nu = pin(st["twists"]), pin(st["poses"]), pin(st["null_poses"])
h_out = {"wrench": torch.empty((n, 6), dtype=torch.float64).pin_memory(),
         "autodyn": torch.empty((n, 6), dtype=torch.float64).pin_memory(),
         "ctrl": torch.empty((n, 36), dtype=torch.float64).pin_memory(), "regressor": None}
b = ContinuousContactModelBatch(0)
b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
d_tw, d_po, d_nu = tw.cuda(), po.cuda(), nu.cuda()
d_out = b.alloc_aos_outputs(n, FULL)


def timeit(call, reps=5):
    call()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        call()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def report(name, secs, up, down):
    print(f"{name:64s} {secs*1e3:8.2f} ms  {n/secs/1e6:7.1f} M evals/s  up {up*n/secs/1e9:5.1f} GB/s  down {down*n/secs/1e9:5.1f} GB/s",
          flush=True)


ref = b.evaluate_aos(d_tw, d_po, d_nu, None, FULL)
torch.cuda.synchronize()
for mask, name, down in ((FULL, "full (dense ctrl)", 384), (WRENCH | AUTODYN, "wrench + autodyn", 96)):
    c, _ = b.prepare_aos(tw, po, nu, None, mask, out=h_out)
    report(f"AoS kernel, {name}: everything in pinned host memory", timeit(c), 240, down)
    if mask == FULL:
        for k in ("wrench", "autodyn", "ctrl"):
            assert torch.equal(h_out[k], ref[k].cpu()), k
    c, _ = b.prepare_aos(tw, po, nu, None, mask, out=d_out)
    report(f"AoS kernel, {name}: inputs in host memory, outputs in HBM", timeit(c), 240, 0)
    c, _ = b.prepare_aos(d_tw, d_po, d_nu, None, mask, out=h_out)
    report(f"AoS kernel, {name}: inputs in HBM, outputs in host memory", timeit(c), 0, down)
t = timeit(lambda: b.evaluate_host(tw, po, nu, None, FULL, out=h_out))
report("blf_ccm_eval_batch_host (copy engines, compact ctrl + expansion)", t, 240, 160)
b.set_host_threads(0)
t = timeit(lambda: b.evaluate_host(tw, po, nu, None, FULL, out=h_out))
report("blf_ccm_eval_batch_host (copy engines, dense download)", t, 240, 384)

#!/usr/bin/env python
"""Sweep of the host-buffer evaluation path (blf_ccm_eval_batch_host): worker threads of the compact
control-matrix expansion x chunk size, the dense download it replaces, and (BLF_CCM_TUNE_HOST_NOEXPAND=1
in the environment) the same pipeline with the expansion skipped -- which separates the PCIe part
from the host part.   python tools/host_sweep.py [n]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from bipedal_locomotion_framework_b200 import synthetic as syn
from bipedal_locomotion_framework_b200.contact_models import FULL, ContinuousContactModelBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
st = syn.make_states(n, seed=46)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
tw, po, nu = pin(st["twists"]), pin(st["poses"]), pin(st["null_poses"])
out = {"wrench": torch.empty((n, 6), dtype=torch.float64).pin_memory(),
       "autodyn": torch.empty((n, 6), dtype=torch.float64).pin_memory(),
       "ctrl": torch.empty((n, 36), dtype=torch.float64).pin_memory(), "regressor": None}
b = ContinuousContactModelBatch(0)
b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
print(f"n = {n}, cores = {os.cpu_count()}, NOEXPAND = {os.environ.get('BLF_CCM_TUNE_HOST_NOEXPAND', '0')}")


def rate(reps=6):
    b.evaluate_host(tw, po, nu, None, FULL, out=out)
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        b.evaluate_host(tw, po, nu, None, FULL, out=out)
        best = min(best, time.perf_counter() - t0)
    return n / best / 1e6, best * 1e3


for chunk in (32768, 65536, 131072, 262144):
    b.set_host_chunk(chunk)
    for threads in (0, 1, 2, 4, 6, 8, 12, 16):
        if threads > (os.cpu_count() or 1):
            continue
        b.set_host_threads(threads)
        r, ms = rate()
        print(f"chunk {chunk:7d}  threads {threads:2d} ({'dense download' if threads == 0 else 'compact + expand'}): "
              f"{r:7.1f} M evals/s  {ms:7.2f} ms")

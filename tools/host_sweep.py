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
print(f"n = {n}, cores = {os.cpu_count()}, NOEXPAND = {os.environ.get('BLF_CCM_TUNE_HOST_NOEXPAND', '0')}")
quick = "--quick" in sys.argv


def rate(b, reps=6):
    b.evaluate_host(tw, po, nu, None, FULL, out=out)
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        b.evaluate_host(tw, po, nu, None, FULL, out=out)
        best = min(best, time.perf_counter() - t0)
    return n / best / 1e6, best * 1e3


# the stream / schedule knobs are read when a handle is created
for up, down, ramp in ((1, 1, 0), (1, 1, 1), (2, 1, 1), (3, 1, 1), (2, 2, 1), (2, 1, 0)):
    os.environ["BLF_CCM_TUNE_HOST_UP"], os.environ["BLF_CCM_TUNE_HOST_DOWN"] = str(up), str(down)
    os.environ["BLF_CCM_TUNE_HOST_RAMP"] = str(ramp)
    b = ContinuousContactModelBatch(0)
    b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
    for chunk in ((65536, 131072, 262144) if not quick else (131072,)):
        b.set_host_chunk(chunk)
        for threads in ((0, 2, 4, 6, 8) if not quick else (4,)):
            b.set_host_threads(threads)
            r, ms = rate(b)
            print(f"up {up} down {down} ramp {ramp}  chunk {chunk:7d}  threads {threads:2d} "
                  f"({'dense download' if threads == 0 else 'compact + expand'}): {r:7.1f} M evals/s  {ms:7.2f} ms",
                  flush=True)
    del b

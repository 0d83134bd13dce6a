#!/usr/bin/env python
"""Whole dynamics step (J^T wrench accumulation, then the mass-matrix solve) with and without programmatic
dependent launch between the two kernels, beside the two kernels timed alone."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from bipedal_locomotion_framework_b200 import synthetic as syn
from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch
from bipedal_locomotion_framework_b200.system import FloatingBaseDynamicsBatch, GeneralizedForceBatch


def timeit(fn, iters=30, warm=5):
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


rnd = lambda *s: torch.rand(s, dtype=torch.float64, device="cuda") * 2 - 1
for nopdl in (0, 1):
    os.environ["BLF_CCM_TUNE_NO_PDL"] = str(nopdl)
    b = ContinuousContactModelBatch(0)
    b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
    dyn, gf = FloatingBaseDynamicsBatch(b), GeneralizedForceBatch(b)
    for ns, cps, nc in ((409600, 2, 29), (409600, 2, 12), (409600, 2, 18)):
        n = ns * cps
        st = syn.make_states(min(n, 1 << 18), seed=49)
        reps = (n + st["n"] - 1) // st["n"]
        pl = torch.from_numpy(np.ascontiguousarray(np.tile(
            syn.aos_to_planes(st["twists"], st["poses"], st["null_poses"]), (1, reps))[:, :n])).cuda()
        nb = 3
        Js = [rnd(n, 6, nc) for _ in range(nb)]
        Ms = []
        for _ in range(nb):
            A = rnd(ns, nc, nc)
            M = torch.bmm(A, A.transpose(1, 2)) / nc
            del A
            M.diagonal(dim1=1, dim2=2).add_(0.5)
            Ms.append(M)
        bias, tau = rnd(ns, nc), rnd(ns, nc - 6)
        outs = [torch.empty_like(bias) for _ in range(nb)]
        whole = [dyn.prepare_acceleration(cps, pl, Js[j], bias, Ms[j], tau, out=outs[j])[0] for j in range(nb)]
        solve = [dyn.prepare_solve(Ms[j], bias, tau, out=outs[j])[0] for j in range(nb)]
        force = [gf.prepare(cps, nc, pl, Js[j], bias, out=outs[j])[0] for j in range(nb)]
        tw, ts, tf = (timeit(lambda i, c=c: c[i % nb]()) for c in (whole, solve, force))
        print(f"PDL {'off' if nopdl else 'on '}  nc {nc:2d}: whole step {tw*1e3:7.1f} us   solve alone {ts*1e3:7.1f} us   "
              f"J^T wrench alone {tf*1e3:7.1f} us   sum {1e3*(ts+tf):7.1f} us", flush=True)
        del Js, Ms, outs, whole, solve, force, pl
        torch.cuda.empty_cache()

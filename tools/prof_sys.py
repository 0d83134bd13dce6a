#!/usr/bin/env python
"""Launch a few iterations of one System-row kernel, for ncu (development aid, not the bench).

  python tools/prof_sys.py rollout_small | rollout_large | rollout_large_rho | rollout_full |
                           genforce | genforce_small | genforce_narrow | genforce_6x2 | genforce_12x2 | euler
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from bipedal_locomotion_framework_b200 import synthetic as syn
from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch
from bipedal_locomotion_framework_b200.system import (GeneralizedForceBatch, KinematicsBatch,
                                                      RolloutBatch)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "rollout_small"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    b = ContinuousContactModelBatch(0)
    b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
    rnd = lambda *s: torch.rand(s, dtype=torch.float64, device="cuda") * 2 - 1
    if what.startswith("rollout"):
        nr, feet, H = (4096, 2, 100) if what == "rollout_small" else (65536, 2, 100)
        rho = 2.0 if what.endswith("rho") else 0.0
        mask = 7 if what == "rollout_full" else 0
        chains = nr * feet
        st = syn.make_states(min(chains, 1 << 17), seed=48)
        reps = (chains + st["n"] - 1) // st["n"]
        tile = lambda a: torch.from_numpy(np.ascontiguousarray(np.tile(a.T, (1, reps))[:, :chains])).cuda()
        pos, rot, null = tile(st["poses"][:, :3]), tile(st["poses"][:, 3:]), tile(st["null_poses"])
        tw = rnd(6, H * chains)
        call, out = RolloutBatch(b).prepare(nr, feet, H, 0.01, rho, tw, pos, rot, null,
                                            [0, 0, 30., 0, 0, 0], [1., 10.], mask=mask)
    elif what.startswith("genforce"):
        ns, cps, ncols = {"genforce_small": (409600, 2, 29), "genforce_narrow": (1 << 21, 1, 6),
                          "genforce_6x2": (1 << 21, 2, 6), "genforce_12x2": (1 << 20, 2, 12)}.get(
            what, (1 << 21, 2, 29))
        n = ns * cps
        st = syn.make_states(min(n, 1 << 17), seed=49)
        reps = (n + st["n"] - 1) // st["n"]
        pl = torch.from_numpy(np.ascontiguousarray(np.tile(
            syn.aos_to_planes(st["twists"], st["poses"], st["null_poses"]), (1, reps))[:, :n])).cuda()
        J, base = rnd(n, 6, ncols), rnd(ns, ncols)
        call, out, _ = GeneralizedForceBatch(b).prepare(cps, ncols, pl, J, base)
    else:
        n = 1 << 23
        st = syn.make_states(1 << 17, seed=47)
        rot = torch.from_numpy(np.ascontiguousarray(np.tile(st["poses"][:, 3:].T, (1, n >> 17)))).cuda()
        call = KinematicsBatch(0, b.handle).prepare_euler_step(2.0, 1e-4, rnd(6, n), rnd(3, n), rot)
    for _ in range(iters):
        call()
    torch.cuda.synchronize()
    print("ok", what)


if __name__ == "__main__":
    main()

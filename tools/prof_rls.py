#!/usr/bin/env python
"""ncu target: 20 launches of the RLS batch update (p = 2, m = 6) at 8.4 M estimators through the
pre-bound call.  BLF_CCM_TUNE_RLS_PIPE=1 selects the plain kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bipedal_locomotion_framework_b200 import synthetic as syn
from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch
from bipedal_locomotion_framework_b200.estimators import RecursiveLeastSquareBatch
n = 1 << 23
b = ContinuousContactModelBatch(0); b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
rls = RecursiveLeastSquareBatch(b, [1.0, 1.0, 1.0, 0.01, 0.01, 0.01], 0.99)
th = torch.rand((2, n), dtype=torch.float64, device="cuda") * 1e3 + 10
P = torch.zeros((4, n), dtype=torch.float64, device="cuda"); P[0] = 1e6; P[3] = 1e4
zz = torch.randn((6, n), dtype=torch.float64, device="cuda")
Yp = torch.randn((12, n), dtype=torch.float64, device="cuda") * 1e-2
call = rls.prepare_advance(Yp, zz, th, P)
for _ in range(20):
    call()
torch.cuda.synchronize()
print("ok")

#!/usr/bin/env python
"""blf_ccm_rollout_integrate_cost_host (configs[2] shape from pinned host memory) under the time-chunk
size knob BLF_CCM_TUNE_ROLLOUT_CHUNK_MB."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "one":
    import numpy as np, torch
    from bipedal_locomotion_framework_b200 import synthetic as syn
    from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch
    from bipedal_locomotion_framework_b200.system import RolloutBatch
    S, F, H = 4096, 2, 100
    chains, n = S * F, S * F * H
    st = syn.make_states(n, seed=45)
    planes = syn.aos_to_planes(st["twists"], st["poses"], st["null_poses"])
    b = ContinuousContactModelBatch(0); b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
    rb = RolloutBatch(b)
    hp = torch.from_numpy(np.ascontiguousarray(planes)).pin_memory()
    args = (S, F, H, 0.01, 0.01, hp[0:6], hp[6:9, :chains], hp[9:18, :chains], hp[18:30, :chains],
            [0, 0, 30.0, 0, 0, 0], [1.0, 10.0])
    for _ in range(5):
        r = rb.run_host(*args, want_cost=False)
    K, best = 40, 1e9
    t0 = time.perf_counter()
    for _ in range(K):
        t1 = time.perf_counter(); r = rb.run_host(*args, want_cost=False); best = min(best, time.perf_counter() - t1)
    mean = (time.perf_counter() - t0) / K
    print(f"chunk_mb={os.environ.get('BLF_CCM_TUNE_ROLLOUT_CHUNK_MB','8(default)')}: mean {mean*1e3:.3f} ms "
          f"({n/mean/1e6:.0f} M evals/s, {(n*48+chains*192)/mean/1e9:.1f} GB/s up), best {best*1e3:.3f} ms, argmin {r[1]}")
else:
    for mb in ("2", "4", "8", "16", "32", "64"):
        subprocess.run([sys.executable, __file__, "one"], env=dict(os.environ, BLF_CCM_TUNE_ROLLOUT_CHUNK_MB=mb))

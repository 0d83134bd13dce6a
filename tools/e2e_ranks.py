#!/usr/bin/env python
"""Host-buffer evaluation (blf_ccm_eval_batch_host) on all ranks of one box at once, over expansion
thread counts and chunk sizes: what the 8-GPU e2e figure is made of.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_ranks.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from bipedal_locomotion_framework_b200 import synthetic as syn
from bipedal_locomotion_framework_b200.contact_models import FULL, ContinuousContactModelBatch

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo", rank=rank, world_size=world)
n = 1 << 21
st = syn.make_states(n, seed=46, start=rank * n)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
tw, po, nu = pin(st["twists"]), pin(st["poses"]), pin(st["null_poses"])
out = {"wrench": torch.empty((n, 6), dtype=torch.float64).pin_memory(),
       "autodyn": torch.empty((n, 6), dtype=torch.float64).pin_memory(),
       "ctrl": torch.empty((n, 36), dtype=torch.float64).pin_memory(), "regressor": None}
b = ContinuousContactModelBatch(local)
b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
if rank == 0:
    print(f"ranks {world}, cpus {os.cpu_count()}, n per rank {n}", flush=True)
for chunk in (131072, 262144):
    b.set_host_chunk(chunk)
    for threads in (0, 1, 2, 3, 4, 6):
        b.set_host_threads(threads)
        b.evaluate_host(tw, po, nu, None, FULL, out=out)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(6):
            b.evaluate_host(tw, po, nu, None, FULL, out=out)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"chunk {chunk:7d} threads {threads} ({'dense download' if threads == 0 else 'compact + expand'}): "
                  f"{world * n * 6 / dt.item() / 1e6:8.1f} M evals/s over {world} ranks", flush=True)
if world > 1:
    dist.destroy_process_group()

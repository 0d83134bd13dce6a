#!/usr/bin/env python
"""Measured PCIe ceiling of the box (pinned host memory, 256 MiB transfers): H2D alone, D2H alone,
both directions at once, and both at once in the e2e path's 240 : 384 byte ratio.  Development aid:
says how far blf_ccm_eval_batch_host (the e2e number) is from what the link can do."""
import torch

MB = 1 << 20
dev = torch.device("cuda", 0)
h_in = torch.empty(256 * MB, dtype=torch.uint8).pin_memory()
h_out = torch.empty(410 * MB, dtype=torch.uint8).pin_memory()
d_in = torch.empty(256 * MB, dtype=torch.uint8, device=dev)
d_out = torch.empty(410 * MB, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def h2d(nbytes=256 * MB):
    with torch.cuda.stream(s1):
        d_in[:nbytes].copy_(h_in[:nbytes], non_blocking=True)


def d2h(nbytes=256 * MB):
    with torch.cuda.stream(s2):
        h_out[:nbytes].copy_(d_out[:nbytes], non_blocking=True)


def both(a=256 * MB, b=256 * MB):
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    h2d(a)
    d2h(b)


t = timed(lambda: (s1.wait_stream(torch.cuda.current_stream()), h2d()))
print(f"H2D alone            {256 * MB / t / 1e9:6.1f} GB/s")
t = timed(lambda: (s2.wait_stream(torch.cuda.current_stream()), d2h()))
print(f"D2H alone            {256 * MB / t / 1e9:6.1f} GB/s")
t = timed(both)
print(f"both, 1:1            {256 * MB / t / 1e9:6.1f} GB/s each direction")
a, b = 240 * MB, 384 * MB
t = timed(lambda: both(a, b))
print(f"both, 240:384        H2D {a / t / 1e9:6.1f} GB/s  D2H {b / t / 1e9:6.1f} GB/s  -> "
      f"{MB / t / 1e6:6.1f} M evals/s ceiling for 240 B in + 384 B out per evaluation")

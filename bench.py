#!/usr/bin/env python
"""bench.py -- contact evals/sec (wrench + 6x6 Jacobians, FP64) of the batched
ContinuousContactModel evaluation, with roofline, CPU baseline and end-to-end numbers.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch of synthetic contact states.  Workload at every
N: BASELINE.json configs[2] -- the sampling-MPC rollout batch, 2 feet x 4096 samples x 100 horizon
steps = 819 200 states per GPU, wrench + autonomous dynamics + control matrix, uniform foot, SoA
planes in / SoA wrench+autodyn planes and dense 6x6 out (600 algorithmic bytes per evaluation).
N > 1 is weak scaling: every rank owns its own 4096 samples (rollouts), no data-path collective in
the evaluation; the sampling-MPC epilogue (per-rollout cost, arg-min, NCCL all-gather of one
16-byte pair per rank) is timed separately and reported under "mpc".

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FEET, SAMPLES, HORIZON = 2, 4096, 100
N_PER_GPU = FEET * SAMPLES * HORIZON          # 819 200
ROLLOUT_LEN = FEET * HORIZON                  # 200 evaluations per rollout
BYTES_FULL_UNIFORM = 600                      # 27 live input doubles + 48 output doubles
METRIC = "contact evals/sec (wrench+6x6 Jacobians, FP64)"
UNIT = "evals/s"
WORKLOAD = ("configs[2]: sampling-MPC rollout batch 2 feet x 4096 samples x 100 steps per GPU, "
            "wrench+autodyn+ctrl, uniform params, SoA")


_JSON_OUT = None


def emit(line: dict) -> None:
    """The one JSON line, on the process's original stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the bench runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting",
               0x100: "display_clock_setting", 0x10: "sync_boost"}

    def __init__(self, index: int, period: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples = []
        self.stop_flag = False
        self.ok = False
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), clk, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self, t0: float, t1: float) -> dict:
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed"
        if len(inside) < 3:
            inside, window = list(self.samples), "warmup+timed (timed region shorter than 3 samples)"
        clks = sorted(s[1] for s in inside) or [0]
        bits = 0
        for s in inside:
            bits |= s[2]
        reasons = [name for bit, name in self.REASONS.items() if bits & bit]
        return {"sm_mhz": clks[len(clks) // 2], "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(inside), "window": window}


# --------------------------------------------------------------------------------------------------
# CPU arm (oracle): bench.py may execute oracle/ only here
# --------------------------------------------------------------------------------------------------

def _cpu_impl(prefer_reference: bool = True):
    """(module, kind, what): oracle/_ref (the reference's own sources compiled against stand-in
    Eigen/iDynTree headers, kind "reference") when it was built, else the C port (kind "port")."""
    from oracle import ccm_oracle, ref_binding
    if prefer_reference and ref_binding.usable():
        return ref_binding, "reference", ("oracle/_ref: the reference's ContinuousContactModel.cpp compiled in place "
                                          "against stand-in Eigen/iDynTree headers (eager evaluation), g++ -O3 -DNDEBUG = CMake Release")
    ccm_oracle.build()
    return ccm_oracle, "port", "oracle/ccm_oracle.c per-instance path, gcc -O2 -ffp-contract=off"


def cpu_reference_pass(st, nthreads: int, min_seconds: float, mask: int = 7, prefer_reference: bool = True):
    """Time the reference's per-instance path (threaded) on the given states."""
    impl, kind, what = _cpu_impl(prefer_reference)
    n = st["twists"].shape[0]
    best = None
    spent = 0.0
    passes = 0
    while spent < min_seconds or passes < 2:
        t0 = time.perf_counter()
        impl.eval_batch_states(st, mask=mask, nthreads=nthreads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        spent += dt
        passes += 1
    return n / best, passes, spent, kind, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from bipedal_locomotion_framework_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    sample_n = N_PER_GPU  # one full per-GPU batch per step
    st = syn.make_states(sample_n, seed=42 + 3)
    impl, kind, what = _cpu_impl()
    for _ in range(max(1, min(args.warmup, 3))):
        impl.eval_batch_states(st, mask=7, nthreads=cores)
    steps = max(1, min(args.steps, 50))  # bounded: each step is one full 819 200-state pass
    t0 = time.perf_counter()
    for _ in range(steps):
        impl.eval_batch_states(st, mask=7, nthreads=cores)
    dt = time.perf_counter() - t0
    value = sample_n * steps / dt
    sample = (f"{steps} passes over one {sample_n}-state batch (configs[2] per-GPU shard), {cores} threads, "
              f"one model object per thread: setState, setNullForceTransform, getContactWrench, "
              f"getAutonomousDynamics, getControlMatrix per state; {what}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "evals_per_step": sample_n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": ("the reference's own sources, compiled from /root/reference into oracle/_ref against stand-in "
                 "Eigen/iDynTree headers (the real libraries are absent from the image)" if kind == "reference"
                 else "oracle/_ref was not built (no /root/reference at build time): C port timed"),
    }
    if kind == "reference":  # the leaner C restatement beside it, for scale
        v, passes, spent, _, pwhat = cpu_reference_pass(st, cores, 2.0, prefer_reference=False)
        line["cpu_baseline"]["port_value"] = v
        line["cpu_baseline"]["port_sample"] = f"best of {passes} passes, {pwhat}"
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist

    from bipedal_locomotion_framework_b200 import sharding
    from bipedal_locomotion_framework_b200 import synthetic as syn
    from bipedal_locomotion_framework_b200.contact_models import FULL, ContinuousContactModelBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the contact-model backend has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    K, Wm = args.steps, max(args.warmup, 3)
    n = N_PER_GPU
    batch = ContinuousContactModelBatch(local)
    batch.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)

    # weak scaling: world*4096 rollouts in total, block-partitioned; rank r owns rollouts
    # [first, first+count) = states first*200 .. of the seeded stream
    first_rollout, n_roll = sharding.shard_rollouts(world * SAMPLES, world, rank)
    assert n_roll * ROLLOUT_LEN == n
    st = syn.make_states(n, seed=42 + 3, start=first_rollout * ROLLOUT_LEN)
    planes_np = syn.aos_to_planes(st["twists"], st["poses"], st["null_poses"])
    NSETS = 3   # rotate distinct input/output buffer sets so no step finds its data in L2
    planes = [torch.from_numpy(planes_np).to(dev) for _ in range(NSETS)]
    outs = [batch.alloc_soa_outputs(n, FULL) for _ in range(NSETS)]
    calls = [batch.prepare_soa(planes[j], None, FULL, out=outs[j])[0] for j in range(NSETS)]
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()

    def step(i):
        calls[i % NSETS]()   # one blf_ccm_eval_batch_soa call (arguments bound once)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(Wm):
        step(i)
    barrier()
    launches0 = batch.handle.launch_count
    t_host0 = time.perf_counter()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for i in range(K):
        step(i)
    e_end.record()
    barrier()
    t_host1 = time.perf_counter()
    launches = batch.handle.launch_count - launches0
    total_ms = e_start.elapsed_time(e_end)
    # the timed region is K back-to-back launches of one kernel on one stream, so its average
    # launch duration (gaps included) is region / K
    avg_kernel_ms = total_ms / K
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * n * K / (total_ms * 1e-3)
    # per-launch spread, outside the timed region (event pairs perturb the loop slightly)
    Kp = min(K, 200)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(Kp)]
    for i in range(Kp):
        ev[i][0].record()
        step(i)
        ev[i][1].record()
    torch.cuda.synchronize()
    kern_ms = sorted(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.summary(t_host0, t_host1)

    if args.only_main:
        sampler.stop_flag = True
        if rank == 0:
            emit({"only_main": True, "value": value, "ms_per_step": total_ms / K,
                  "avg_launch_ms": avg_kernel_ms, "gpu_launches": int(launches)})
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- sampling-MPC epilogue: evaluate + per-rollout cost + arg-min (+ NCCL all-gather) --------
    ref_wrench, weights = [0.0, 0.0, 30.0, 0.0, 0.0, 0.0], [1.0, 10.0]
    mpc_calls = [batch.prepare_rollout(planes[j], ROLLOUT_LEN, ref_wrench, weights, mask=FULL,
                                       index_base=first_rollout, out=outs[j], want_cost=False)
                 for j in range(NSETS)]

    # the only exchange of the path: every rank's 16-byte (cost, index) pair.  Default: stores into
    # the peers' mailboxes over NVLink (one single-warp kernel per rank); NCCL all-gather as the
    # baseline it replaces (also timed below)
    peer, peer_note = None, None
    if world > 1 and not args.nccl_argmin:
        try:
            peer = sharding.PeerArgmin(batch, world, rank, dist)
        except Exception as e:   # e.g. CUDA IPC not permitted in this container
            peer_note = f"peer-memory mailbox unavailable ({e}); NCCL all-gather used"
    if world > 1:   # the choice must be the same on every rank
        flag = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            peer = None

    def exchange(best, use_peer=True):
        if world == 1:
            return best
        if peer is not None and use_peer:
            return peer.global_best      # written by the rollout's own reduction kernel (fused)
        return batch.argmin_pairs(sharding.all_gather_pairs(best, world, dist))

    if peer is not None:
        peer.fuse_into_rollouts(True)

    def mpc_step(i, use_peer=True):
        call, _, _, best = mpc_calls[i % NSETS]
        call()
        return exchange(best, use_peer)

    for i in range(Wm):
        gbest = mpc_step(i)
    barrier()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Km = min(K, 200)
    m0.record()
    for i in range(Km):
        gbest = mpc_step(i)
    m1.record()
    barrier()
    mpc_ms = m0.elapsed_time(m1)
    if world > 1:
        t = torch.tensor([mpc_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mpc_ms = float(t.item())
    best_cost, best_idx = batch.decode_best(gbest)
    nccl_ms = None
    if world > 1 and peer is not None:   # the same step with the NCCL all-gather it replaces
        peer.fuse_into_rollouts(False)
        for i in range(Wm):
            mpc_step(i, use_peer=False)
        barrier()
        m0.record()
        for i in range(Km):
            mpc_step(i, use_peer=False)
        m1.record()
        barrier()
        t = torch.tensor([m0.elapsed_time(m1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nccl_ms = float(t.item()) / Km
        peer.fuse_into_rollouts(True)
    collective = "none (1 GPU)"
    if world > 1:
        collective = ("peer-memory mailbox over NVLink, fused into the cost-reduction kernel "
                      "(blf_ccm_rollout_set_exchange): no extra launch, no collective library" if peer is not None else "nccl all_gather 16 B/rank + device arg-min")
    mpc = {"value": world * n * Km / (mpc_ms * 1e-3), "unit": UNIT, "ms_per_step": mpc_ms / Km,
           "steps": Km, "rollouts": world * n_roll, "rollout_len": ROLLOUT_LEN,
           "argmin": {"cost": best_cost, "rollout": best_idx},
           "collective": collective, "ms_per_step_with_nccl_all_gather": nccl_ms, "note": peer_note,
           "hbm_frac_of_measured": None}

    # ---- fused rollout: integrate -> contact model -> cost, pose in ---------
    # registers; same shape (2 feet x 4096 samples x 100 steps), the twist planes of the batch read
    # time-major, initial / null poses = the first 8192 states; Baumgarte rho = reference default
    from bipedal_locomotion_framework_b200.system import GeneralizedForceBatch, KinematicsBatch, RolloutBatch
    rb = RolloutBatch(batch)
    chains = FEET * SAMPLES
    fused_calls = [rb.prepare(SAMPLES, FEET, HORIZON, 0.01, 0.01, planes[j][0:6],
                              planes[j][6:9, :chains], planes[j][9:18, :chains],
                              planes[j][18:30, :chains], ref_wrench, weights, mask=0,
                              index_base=first_rollout, want_cost=False) for j in range(NSETS)]

    def fused_step(i):
        call, o = fused_calls[i % NSETS]
        call()
        return exchange(o["best"])

    for i in range(Wm):
        fbest = fused_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(Km):
        fbest = fused_step(i)
    f1.record()
    barrier()
    fused_ms = f0.elapsed_time(f1)
    if world > 1:
        t = torch.tensor([fused_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        fused_ms = float(t.item())
    fc, fi = batch.decode_best(fbest)
    mpc_fused = {"value": world * n * Km / (fused_ms * 1e-3), "unit": UNIT,
                 "ms_per_step": fused_ms / Km, "steps": Km,
                 "what": "blf_ccm_rollout_integrate_cost: ForwardEuler(FloatingBaseSystemKinematics) "
                         "-> contact wrench -> cost -> arg-min, cost only; 48 B (twist) per "
                         "evaluation from HBM, pose in registers",
                 "rho": 0.01, "dT": 0.01, "argmin": {"cost": fc, "rollout": fi},
                 "speedup_vs_unfused_mpc": (mpc_ms / Km) / (fused_ms / Km)}

    # end to end for the fused rollout: twists and poses in pinned HOST memory, one pair back
    if world == 1:
        hp = torch.from_numpy(planes_np).pin_memory()
        host_args = (SAMPLES, FEET, HORIZON, 0.01, 0.01, hp[0:6], hp[6:9, :chains], hp[9:18, :chains],
                     hp[18:30, :chains], ref_wrench, weights)
        for _ in range(3):
            rb.run_host(*host_args, want_cost=False)
        Kh = max(3, min(K, 30))
        t0 = time.perf_counter()
        for _ in range(Kh):
            hb = rb.run_host(*host_args, want_cost=False)
        host_s = (time.perf_counter() - t0) / Kh
        mpc_fused["e2e"] = {"value": n / host_s, "unit": UNIT, "ms_per_step": host_s * 1e3, "steps": Kh,
                            "h2d_bytes_per_step": int(n * 48 + chains * 24 * 8),
                            "d2h_bytes_per_step": 16,
                            "api": "blf_ccm_rollout_integrate_cost_host (time-major twist planes and "
                                   "per-chain poses in pinned host memory in, arg-min pair out)",
                            "argmin": {"cost": hb[0], "rollout": hb[1]}}
        if not args.no_cpu:
            from oracle import sys_oracle
            cores = os.cpu_count() or 1
            best_s = None
            for _ in range(3):
                t0 = time.perf_counter()
                sys_oracle.rollout(SAMPLES, FEET, HORIZON, 0.01, 0.01, planes_np[0:6],
                                   planes_np[6:9, :chains], planes_np[9:18, :chains],
                                   planes_np[18:30, :chains], uniform=syn.REFERENCE_TEST_PARAMS, mask=0,
                                   wrench_ref=ref_wrench, weights=weights, nthreads=cores)
                dt_ = time.perf_counter() - t0
                best_s = dt_ if best_s is None else min(best_s, dt_)
            mpc_fused["cpu_baseline"] = {"value": n / best_s, "unit": UNIT, "cores": cores, "kind": "port",
                                         "sample": "best of 3 passes of oracle/sys_oracle.c rollout over "
                                                   "the same 4096 x 2 x 100 batch"}

    # ---- the other rows next to the path, one GPU, briefly (fractions of the measured HBM peak) ---
    next_rows = None
    if world == 1:
        def timed(fn, iters=50):
            for i in range(5):
                fn(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(iters):
                fn(i)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / iters
        gf = GeneralizedForceBatch(batch)
        ncols, cps = 29, FEET                        # 6 + 23 DoF (iCub-sized), two feet per robot
        Js = [torch.rand((n, 6, ncols), dtype=torch.float64, device=dev) for _ in range(2)]
        base = torch.rand((n // cps, ncols), dtype=torch.float64, device=dev)
        gouts = [torch.empty_like(base) for _ in range(2)]
        gcalls = [gf.prepare(cps, ncols, planes[j], Js[j], base, out=gouts[j])[0] for j in range(2)]
        g_ms = timed(lambda i: gcalls[i % 2]())
        g_bytes = n * (200 + 48 * ncols) + (n // cps) * ncols * 16
        kb = KinematicsBatch(local, batch.handle)
        kp = [planes[j][6:18].clone() for j in range(NSETS)]
        kcalls = [kb.prepare_euler_step(0.01, 1e-4, planes[j][0:6], kp[j][0:3], kp[j][3:12])
                  for j in range(NSETS)]
        k_ms = timed(lambda i: kcalls[i % NSETS]())
        pk = _peaks()[0]
        next_rows = {
            "generalized_force": {"what": "blf_ccm_generalized_force_soa: base + sum J^T wrench, "
                                          f"{n // cps} systems x {cps} contacts, 6 x {ncols} Jacobians",
                                  "ms": g_ms, "contacts_per_s": n / (g_ms * 1e-3),
                                  "algorithmic_gbs": g_bytes / (g_ms * 1e-3) / 1e9,
                                  "hbm_frac_of_measured": g_bytes / (g_ms * 1e-3) / 1e9 / pk},
            "kinematics_euler_step": {"what": "blf_sys_kinematics_euler_step_soa, rho 0.01, 240 B/system",
                                      "ms": k_ms, "systems_per_s": n / (k_ms * 1e-3),
                                      "hbm_frac_of_measured": n * 240 / (k_ms * 1e-3) / 1e9 / pk},
        }
        del Js, base, gouts, gcalls, kp, kcalls

    # ---- end to end through the C ABI with HOST buffers (copies inside the timed region) ---------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_tw, h_po, h_nu = pin(st["twists"]), pin(st["poses"]), pin(st["null_poses"])
    h_out = {"wrench": torch.empty((n, 6), dtype=torch.float64).pin_memory(),
             "autodyn": torch.empty((n, 6), dtype=torch.float64).pin_memory(),
             "ctrl": torch.empty((n, 36), dtype=torch.float64).pin_memory(), "regressor": None}
    Ke = max(3, min(K, 30))
    for _ in range(3):
        batch.evaluate_host(h_tw, h_po, h_nu, None, FULL, out=h_out)
    barrier()
    l0 = batch.handle.launch_count
    t0 = time.perf_counter()
    for _ in range(Ke):
        batch.evaluate_host(h_tw, h_po, h_nu, None, FULL, out=h_out)
    e2e_s = time.perf_counter() - t0
    e2e_launches = batch.handle.launch_count - l0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * n * Ke / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(n * 30 * 8), "d2h_bytes_per_step": int(n * 48 * 8),
           "ms_per_step": e2e_s / Ke * 1e3, "steps": Ke, "launches_per_step": e2e_launches / Ke,
           "api": "blf_ccm_eval_batch_host (AoS iDynTree-layout arrays in pinned host memory in/out)"}
    if world == 1:
        # what bounds e2e: the PCIe link, measured here with plain pinned copies of the same byte
        # counts in both directions at once (240 B up, 384 B down per evaluation), no kernel
        d_up = torch.empty(n * 30, dtype=torch.float64, device=dev)
        d_dn = torch.empty(n * 48, dtype=torch.float64, device=dev)
        h_up = torch.empty(n * 30, dtype=torch.float64).pin_memory()
        h_dn = torch.empty(n * 48, dtype=torch.float64).pin_memory()
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

        def both():
            with torch.cuda.stream(s_up):
                d_up.copy_(h_up, non_blocking=True)
            with torch.cuda.stream(s_dn):
                h_dn.copy_(d_dn, non_blocking=True)
        for _ in range(2):
            both()
        torch.cuda.synchronize()
        tp = time.perf_counter()
        for _ in range(5):
            both()
        torch.cuda.synchronize()
        link_s = (time.perf_counter() - tp) / 5
        e2e["pcie"] = {"what": "pinned cudaMemcpyAsync of one step's bytes in both directions at once, no kernel",
                       "h2d_gbs": n * 240 / link_s / 1e9, "d2h_gbs": n * 384 / link_s / 1e9,
                       "ceiling_evals_per_s": n / link_s, "frac_of_ceiling": (n * Ke / e2e_s) / (n / link_s)}
        del d_up, d_dn, h_up, h_dn
    sampler.stop_flag = True

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = _peaks()
    achieved = BYTES_FULL_UNIFORM * n / (avg_kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": _traffic(),
                "kernel": "ccm_soa_kernel<WRENCH|AUTODYN|CTRL, uniform> (1 contact/lane, 1 tile/warp)",
                "algorithmic_bytes_per_launch": BYTES_FULL_UNIFORM * n,
                "avg_launch_ms": avg_kernel_ms,
                "per_launch_event_pairs_ms": {"median": kern_ms[len(kern_ms) // 2], "best": kern_ms[0],
                                              "n": len(kern_ms)},
                "peak_source": peak_src}
    mpc["hbm_frac_of_measured"] = (BYTES_FULL_UNIFORM * n + 8 * n_roll) / (mpc_ms / Km * 1e-3) / 1e9 / peak \
        if world == 1 else None

    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v, passes, spent, kind, what = cpu_reference_pass(st, cores, min_seconds=3.0)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"best of {passes} passes over the same {n}-state batch "
                                  f"({spent:.1f} s wall x {cores} threads), per-instance path; {what}"}
        if kind == "reference":
            pv, pp, _, _, pwhat = cpu_reference_pass(st, cores, 2.0, prefer_reference=False)
            cpu_baseline["port_value"] = pv
            cpu_baseline["port_sample"] = f"best of {pp} passes, {pwhat}"

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "evals_per_step_per_gpu": n, "evals_per_step": world * n,
                   "layout": "SoA planes in; SoA wrench/autodyn planes + dense row-major 6x6 out",
                   "parallelism": f"rollout-sharded x{world}, no data-path collective",
                   "l2": f"{NSETS} rotating input/output buffer sets; one step streams 491.5 MB "
                         "(> 126 MB L2), no explicit flush"},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "mpc": mpc, "mpc_fused": mpc_fused, "next_rows": next_rows,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------------------
# BASELINE.json configs[3] and configs[4]: strong-scaling workloads (run on request)
# --------------------------------------------------------------------------------------------------

def run_large(args):
    import torch
    import torch.distributed as dist

    from bipedal_locomotion_framework_b200 import sharding
    from bipedal_locomotion_framework_b200 import synthetic as syn
    from bipedal_locomotion_framework_b200.contact_models import FULL, WRENCH, ContinuousContactModelBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the contact-model backend has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, Wm = min(args.steps, 50), max(min(args.warmup, 10), 3)
    mask, nsets = FULL, 1

    batch = ContinuousContactModelBatch(local)
    batch.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
    if args.workload == "config4":
        total_rollouts = 335544                   # 335 544 x 200 = 67 108 800 ~ 2^26 states
        first, count = sharding.shard_rollouts(total_rollouts, world, rank)
        n, het, bytes_per = count * ROLLOUT_LEN, True, 632
        total = total_rollouts * ROLLOUT_LEN
        name = ("configs[3]: 64M heterogeneous contact states (per-contact length/width/spring/"
                "damper), rollouts of 200 sharded by rollout, wrench+autodyn+ctrl + per-rollout "
                "cost + NCCL arg-min")
    elif args.workload == "config2":
        total = 1 << 20                           # per GPU (weak): the configuration is a 1-GPU one
        n, het, bytes_per, mask, nsets = total, False, 248, WRENCH, 3
        total *= world
        K = min(args.steps, 2000)
        name = ("configs[1]: 1M random contact states per GPU, wrench only, uniform foot geometry and "
                "stiffness/damping, SoA (25 live planes in, 6 planes out = 248 B/eval)")
    else:
        total = 1 << 28
        first, count = sharding.shard_rollouts(total, world, rank)   # plain block partition
        n, het, bytes_per = count, False, 600
        name = "configs[4]: strong scaling, 256M contact states, wrench+autodyn+ctrl, uniform params"
    planes, prm = syn.make_planes_torch(n, dev, seed=42 + 4 + 1000 * rank, heterogeneous=het)
    out = batch.alloc_soa_outputs(n, mask)
    torch.cuda.synchronize()

    if args.workload == "config4":
        call, _, _, best = batch.prepare_rollout(planes, ROLLOUT_LEN, [0.0, 0.0, 30.0, 0.0, 0.0, 0.0],
                                                 [1.0, 10.0], param_planes=prm, mask=FULL,
                                                 index_base=first, out=out, want_cost=False)

        peer = None
        if world > 1 and not args.nccl_argmin:
            try:
                peer = sharding.PeerArgmin(batch, world, rank, dist)
            except Exception as e:
                print(f"bench.py: peer-memory mailbox unavailable ({e}); NCCL all-gather used", file=sys.stderr)
        if world > 1:   # the choice must be the same on every rank
            flag = torch.tensor([1 if peer is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                peer = None
        if peer is not None:
            peer.fuse_into_rollouts(True)   # the exchange runs inside the cost-reduction kernel

        def step():
            call()
            if world > 1:
                if peer is not None:
                    return peer.global_best
                return batch.argmin_pairs(sharding.all_gather_pairs(best, world, dist))
            return best
    elif nsets > 1:
        # rotating buffer sets: a step never finds its 260 MB in the 126 MB L2
        peer = None
        sets = [(planes, out)] + [([None if q is None else q.clone() for q in planes],
                                  batch.alloc_soa_outputs(n, mask)) for _ in range(nsets - 1)]
        calls = [batch.prepare_soa(pl, None, mask, out=o)[0] for pl, o in sets]
        counter = [0]

        def step():
            calls[counter[0] % nsets]()
            counter[0] += 1
            return None
    else:
        peer = None
        call, _ = batch.prepare_soa(planes, None, FULL, out=out)

        def step():
            call()
            return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(Wm):
        res = step()
    barrier()
    l0 = batch.handle.launch_count
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        res = step()
    e1.record()
    barrier()
    t1 = time.perf_counter()
    launches = batch.handle.launch_count - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.summary(t0, t1)
    sampler.stop_flag = True

    # parity on the same device bits: every state up to 2M, else every 4099th state of this rank's shard
    parity = None
    if rank == 0:
        from oracle import ccm_oracle
        idx = torch.arange(0, n, 1 if n <= (1 << 21) else 4099, device=dev)
        st = syn.sample_states_from_planes(planes, prm, idx)
        ref = ccm_oracle.eval_batch_states(st, mask=7, nthreads=os.cpu_count() or 1)
        worst = 0.0
        got = {"wrench": torch.stack([p[idx] for p in out["wrench"]], 1).cpu().numpy()}
        if mask == FULL:
            got["autodyn"] = torch.stack([p[idx] for p in out["autodyn"]], 1).cpu().numpy()
            got["ctrl"] = out["ctrl"][idx].cpu().numpy()
        for key, blocks in (("wrench", [slice(0, 3), slice(3, 6)]), ("autodyn", [slice(0, 3), slice(3, 6)]),
                            ("ctrl", [slice(6 * q + 3 * (q // 3), 6 * q + 3 * (q // 3) + 3) for q in range(6)])):
            if key not in got:
                continue
            for sl in blocks:
                num = np.abs(got[key][:, sl] - ref[key][:, sl]).max(axis=1)
                den = np.maximum(np.abs(ref[key][:, sl]).max(axis=1), 1e-300)
                worst = max(worst, float(np.where(num == 0, 0, num / den).max()))
        parity = {"sampled_states": int(idx.numel()), "worst_block_rel_err": worst, "tol": 1e-12,
                  "ok": bool(worst <= 1e-12)}

    if rank == 0:
        peak, peak_src = _peaks()
        per_gpu_gbs = bytes_per * n / (ms / K * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": total * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak" if args.workload == "config2" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic (device-generated)",
            "config": {"workload": name, "evals_per_step": total, "evals_per_step_per_gpu": n,
                       "parallelism": f"sharded x{world}" + (
                           (", arg-min pair over the NVLink peer-memory mailbox (fused into the reduction kernel)"
                            if peer is not None else ", NCCL all_gather 16 B/rank")
                           if args.workload == "config4" and world > 1 else ""),
                       "l2": ("3 rotating input/output buffer sets; one step streams 260 MB (> 126 MB L2)"
                              if nsets > 1 else "per-GPU working set >> 126 MB L2")},
            "roofline": {"bound": "hbm", "achieved": per_gpu_gbs, "peak": peak, "unit": "GB/s",
                         "frac": per_gpu_gbs / peak, "traffic": None,
                         "note": "whole step (kernel + epilogue launches) on rank 0's shard", "peak_source": peak_src},
            "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches), "clocks": clocks,
            "parity": parity,
        }
        if res is not None:
            c, i = batch.decode_best(res)
            line["argmin"] = {"cost": c, "rollout": i}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=["config2", "config3", "config4", "config5"],
                    help="config2: 1M states wrench-only; "
                         "config3 (default, the headline): MPC batch, weak scaling; config4: 64M "
                         "heterogeneous states + NCCL arg-min, strong; config5: 256M states, strong")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--nccl-argmin", action="store_true",
                    help="N > 1: use the NCCL all-gather for the arg-min pair instead of the "
                         "peer-memory mailbox")
    ap.add_argument("--only-main", action="store_true",
                    help="profiling aid: run only the main timed loop (no mpc/e2e/cpu legs)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: whatever libraries print there (NCCL's version banner
    # when NCCL_DEBUG is set on the box) is sent to stderr at the file-descriptor level
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    if args.workload != "config3":
        return run_large(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

"""Seeded synthetic contact states.

Counter-based: every uniform draw is splitmix64 of (seed, contact index, lane), so any slice
``[start, start+n)`` of the stream can be produced independently on any rank and the same bits go
to the GPU path and to the CPU oracle.  Base seed 42 matches the reference test's generator seed
(src/ContactModels/tests/ContinousContactModelTest.cpp:66).

Layouts follow iDynTree (see include/blf_ccm.h):
  twists     n x 6   linear xyz, angular xyz
  poses      n x 12  position xyz, rotation 3x3 row-major
  null_poses n x 12  same
  params     n x 4   length, width, spring_coeff, damper_coeff
SoA = 30 planes of n doubles in the order v(0-2) w(3-5) p(6-8) R(9-17) p0(18-20) R0(21-29).
"""
from __future__ import annotations

import numpy as np

# reference test values, ContinousContactModelTest.cpp:43-47 (length, width, spring, damper)
REFERENCE_TEST_PARAMS = (0.12, 0.09, 2000.0, 100.0)

_LANES = 64
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _uniform(seed: int, index: np.ndarray, lane: int) -> np.ndarray:
    """u in [0,1) for every contact index, one lane."""
    with np.errstate(over="ignore"):
        key = _splitmix64(np.asarray([seed], dtype=np.uint64))[0]
        ctr = index.astype(np.uint64) * np.uint64(_LANES) + np.uint64(lane)
        x = _splitmix64(_splitmix64(ctr) ^ key)
    return (x >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def _rpy(roll, pitch, yaw):
    """Rz(yaw) Ry(pitch) Rx(roll), as iDynTree::Rotation::RPY; returns (n,3,3)."""
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    R = np.empty(roll.shape + (3, 3))
    R[..., 0, 0] = cy * cp
    R[..., 0, 1] = cy * sp * sr - sy * cr
    R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp
    R[..., 1, 1] = sy * sp * sr + cy * cr
    R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp
    R[..., 2, 1] = cp * sr
    R[..., 2, 2] = cp * cr
    return R


def make_states(n: int, seed: int = 42, heterogeneous: bool = False, start: int = 0) -> dict:
    """Contact states ``start .. start+n-1`` of stream ``seed`` in AoS layout.

    85 % near-flat feet, 10 % uniform SO(3) (exercises R22 < 0 and the abs/sign quirk of the
    reference), 5 % near-flat times (I + 1e-3 U) (non-orthonormal, as Euler-integrated rotations).
    """
    idx = np.arange(start, start + n, dtype=np.uint64)
    lane = iter(range(_LANES))

    def U(lo, hi):
        return lo + (hi - lo) * _uniform(seed, idx, next(lane))

    p = np.stack([U(-0.05, 0.05) for _ in range(3)], axis=1)
    p0 = p + np.stack([U(-0.01, 0.01) for _ in range(3)], axis=1)
    sel = U(0.0, 1.0)
    roll, pitch, yaw = U(-0.3, 0.3), U(-0.3, 0.3), U(-np.pi, np.pi)
    R = _rpy(roll, pitch, yaw)

    # uniform SO(3): normalised Gaussian 4-vector (Box-Muller) -> quaternion -> matrix
    u = [U(0.0, 1.0) for _ in range(4)]
    r1 = np.sqrt(-2.0 * np.log1p(-u[0]))
    r2 = np.sqrt(-2.0 * np.log1p(-u[2]))
    q = np.stack([r1 * np.cos(2 * np.pi * u[1]), r1 * np.sin(2 * np.pi * u[1]),
                  r2 * np.cos(2 * np.pi * u[3]), r2 * np.sin(2 * np.pi * u[3])], axis=1)
    q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-300)
    qw, qx, qy, qz = q.T
    Rq = np.empty((n, 3, 3))
    Rq[:, 0, 0] = 1 - 2 * (qy * qy + qz * qz)
    Rq[:, 0, 1] = 2 * (qx * qy - qz * qw)
    Rq[:, 0, 2] = 2 * (qx * qz + qy * qw)
    Rq[:, 1, 0] = 2 * (qx * qy + qz * qw)
    Rq[:, 1, 1] = 1 - 2 * (qx * qx + qz * qz)
    Rq[:, 1, 2] = 2 * (qy * qz - qx * qw)
    Rq[:, 2, 0] = 2 * (qx * qz - qy * qw)
    Rq[:, 2, 1] = 2 * (qy * qz + qx * qw)
    Rq[:, 2, 2] = 1 - 2 * (qx * qx + qy * qy)

    pert = np.stack([U(-1.0, 1.0) for _ in range(9)], axis=1).reshape(n, 3, 3)
    Rp = R @ (np.eye(3)[None] + 1e-3 * pert)

    full = (sel >= 0.85) & (sel < 0.95)
    nonortho = sel >= 0.95
    R = np.where(full[:, None, None], Rq, R)
    R = np.where(nonortho[:, None, None], Rp, R)

    R0 = _rpy(U(-0.05, 0.05), U(-0.05, 0.05), yaw + U(-0.1, 0.1))
    v = np.stack([U(-1.0, 1.0) for _ in range(3)], axis=1)
    w = np.stack([U(-1.0, 1.0) for _ in range(3)], axis=1)

    out = {
        "n": n,
        "twists": np.ascontiguousarray(np.concatenate([v, w], axis=1)),
        "poses": np.ascontiguousarray(np.concatenate([p, R.reshape(n, 9)], axis=1)),
        "null_poses": np.ascontiguousarray(np.concatenate([p0, R0.reshape(n, 9)], axis=1)),
        "params": None,
        "uniform": REFERENCE_TEST_PARAMS,
    }
    if heterogeneous:
        L = U(0.08, 0.30)
        W = U(0.04, 0.15)
        k = np.exp(U(np.log(1e3), np.log(1e6)))
        b = np.exp(U(np.log(10.0), np.log(1e4)))
        out["params"] = np.ascontiguousarray(np.stack([L, W, k, b], axis=1))
    return out


def aos_to_planes(twists: np.ndarray, poses: np.ndarray, null_poses: np.ndarray) -> np.ndarray:
    """(30, n) contiguous plane array in the C-ABI's plane order."""
    return np.ascontiguousarray(np.concatenate([twists, poses, null_poses], axis=1).T)


def planes_to_aos(planes: np.ndarray):
    a = np.ascontiguousarray(planes.T)
    return (np.ascontiguousarray(a[:, 0:6]), np.ascontiguousarray(a[:, 6:18]),
            np.ascontiguousarray(a[:, 18:30]))


def reference_test_state(v=(0.3, -0.7, 0.2), w=(-0.5, 0.4, 0.9)) -> dict:
    """BASELINE.json config #1: the reference test's pose and parameters
    (ContinousContactModelTest.cpp:35-47) with a fixed twist in place of the unseeded
    setRandom() one (:40-41)."""
    R = _rpy(np.array([-0.15]), np.array([0.2]), np.array([0.1]))[0]
    pose = np.concatenate([[-0.02, 0.01, 0.005], R.reshape(9)])
    null_pose = np.concatenate([[0.0, 0.0, 0.0], np.eye(3).reshape(9)])
    return {
        "n": 1,
        "twists": np.array([list(v) + list(w)], dtype=np.float64),
        "poses": pose[None].copy(),
        "null_poses": null_pose[None].copy(),
        "params": None,
        "uniform": REFERENCE_TEST_PARAMS,
    }


# --------------------------------------------------------------------------------------------------
# Device-side generation for the configurations that do not fit host memory comfortably
# (BASELINE.json configs[3] = 64M heterogeneous, configs[4] = 256M).  Same distributions as
# make_states, drawn with torch's generator (plumbing), chunked so temporaries stay small.  The
# bits differ from the numpy stream; parity at these sizes is checked by copying a sample of the
# SAME device bits back to the host and running the oracle on it.
# --------------------------------------------------------------------------------------------------

def make_planes_torch(n: int, device, seed: int = 42, heterogeneous: bool = False,
                      live_only: bool = True, chunk: int = 1 << 22):
    """Returns (planes, param_planes): lists of 30 (and 4) 1-D float64 CUDA tensors of n doubles;
    planes that cannot affect any output (R0's third column) are None when live_only."""
    import math
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    dead = {23, 26, 29} if live_only else set()
    planes = [None if i in dead else torch.empty(n, dtype=torch.float64, device=device)
              for i in range(30)]
    prm = [torch.empty(n, dtype=torch.float64, device=device) for _ in range(4)] \
        if heterogeneous else None

    def U(m, lo, hi):
        return lo + (hi - lo) * torch.rand(m, dtype=torch.float64, device=device, generator=g)

    def rpy(r, p, y):
        cr, sr, cp, sp, cy, sy = r.cos(), r.sin(), p.cos(), p.sin(), y.cos(), y.sin()
        return [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                -sp, cp * sr, cp * cr]

    for a in range(0, n, chunk):
        m = min(chunk, n - a)
        sl = slice(a, a + m)
        for c in range(3):
            planes[c][sl] = U(m, -1.0, 1.0)           # v
            planes[3 + c][sl] = U(m, -1.0, 1.0)       # w
            p = U(m, -0.05, 0.05)
            planes[6 + c][sl] = p
            planes[18 + c][sl] = p + U(m, -0.01, 0.01)
        sel = U(m, 0.0, 1.0)
        yaw = U(m, -math.pi, math.pi)
        R = rpy(U(m, -0.3, 0.3), U(m, -0.3, 0.3), yaw)
        # uniform SO(3) from a normalised Gaussian quaternion
        q = torch.randn((4, m), dtype=torch.float64, device=device, generator=g)
        q = q / q.norm(dim=0, keepdim=True).clamp_min(1e-300)
        qw, qx, qy, qz = q
        Rq = [1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw),
              2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw),
              2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]
        full = (sel >= 0.85) & (sel < 0.95)
        nonortho = sel >= 0.95
        for e in range(9):
            r = torch.where(full, Rq[e], R[e])
            r = torch.where(nonortho, r * (1.0 + 1e-3 * U(m, -1.0, 1.0)), r)
            planes[9 + e][sl] = r
        R0 = rpy(U(m, -0.05, 0.05), U(m, -0.05, 0.05), yaw + U(m, -0.1, 0.1))
        for e in range(9):
            if planes[21 + e] is not None:
                planes[21 + e][sl] = R0[e]
        if heterogeneous:
            prm[0][sl] = U(m, 0.08, 0.30)
            prm[1][sl] = U(m, 0.04, 0.15)
            prm[2][sl] = torch.exp(U(m, math.log(1e3), math.log(1e6)))
            prm[3][sl] = torch.exp(U(m, math.log(10.0), math.log(1e4)))
        del R, Rq, R0, q, sel, yaw, full, nonortho
    return planes, prm


def sample_states_from_planes(planes, prm, idx) -> dict:
    """Copy the states at device indices `idx` (1-D int64 tensor) back to host AoS arrays, for the
    oracle.  Dead planes read as 0."""
    import torch
    cols = []
    for p in planes:
        cols.append(torch.zeros(idx.numel(), dtype=torch.float64) if p is None else p[idx].cpu())
    a = torch.stack(cols, dim=1).numpy()
    out = {"n": int(idx.numel()), "twists": np.ascontiguousarray(a[:, 0:6]),
           "poses": np.ascontiguousarray(a[:, 6:18]), "null_poses": np.ascontiguousarray(a[:, 18:30]),
           "params": None, "uniform": REFERENCE_TEST_PARAMS}
    if prm is not None:
        out["params"] = np.ascontiguousarray(torch.stack([q[idx].cpu() for q in prm], dim=1).numpy())
    return out


def make_mass_matrices(n_systems: int, ncols: int, seed: int = 7, spread: float = 1.0) -> np.ndarray:
    """Seeded symmetric positive definite (n_systems, ncols, ncols) matrices shaped like free-floating
    mass matrices: M = D (A A^T / ncols + I/2) D, A ~ U(-1,1), D = diag(10^U(-spread, spread)) -- the
    diagonal scaling spreads the "inertias" over 2*spread decades (a humanoid's span kg to 1e-3 kg m^2),
    so the condition number grows like 100^spread; spread = 0 is benign (cond < 4).  Exactly
    symmetric (both triangles hold the same bits)."""
    rng = np.random.default_rng(seed)
    A = rng.uniform(-1.0, 1.0, (n_systems, ncols, ncols))
    M = A @ A.transpose(0, 2, 1) / ncols
    idx = np.arange(ncols)
    M[:, idx, idx] += 0.5
    if spread > 0:
        d = 10.0 ** rng.uniform(-spread, spread, (n_systems, ncols))
        M = M * d[:, :, None] * d[:, None, :]
    lo = np.tril(M)
    return lo + np.tril(M, -1).transpose(0, 2, 1)

"""Exact evaluation of the last step of FloatingBaseDynamicalSystem::dynamics and the golden-fixture
writer for it (src/System/src/FloatingBaseSystemDynamics.cpp:226-243):

    rhs = known;  rhs.tail(nc - 6) += jointTorques;  acc = (M + reg).llt().solve(rhs)

TEST INFRASTRUCTURE ONLY (see sys_oracle.h).  The Cholesky factor is irrational, but the SOLUTION
acc = (M + reg)^-1 rhs is rational in the inputs: with the double inputs taken as exact rationals
(only the lower triangle of M + reg, mirrored, as LLT reads it) Gaussian elimination over
fractions.Fraction gives it exactly; it is rounded ONCE to double.  The fixture also stores the
infinity-norm condition number of every system (exact inverse), which is what scales the forward
error any floating-point LLT may show.

Run:  python oracle/exact_golden_dyn.py   -> rewrites tests/golden/dyn_exact_golden.npz
"""
from __future__ import annotations

import os
import sys
from fractions import Fraction as F

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
sys.path.insert(0, _ROOT)


def _solve_exact(A, B):
    """A (n x n Fractions, nonsingular), B (n x m) -> A^-1 B, Gauss-Jordan."""
    n, m = len(A), len(B[0])
    W = [list(A[i]) + list(B[i]) for i in range(n)]
    for c in range(n):
        p = next(r for r in range(c, n) if W[r][c] != 0)
        W[c], W[p] = W[p], W[c]
        inv = 1 / W[c][c]
        W[c] = [v * inv for v in W[c]]
        for r in range(n):
            if r != c and W[r][c] != 0:
                f = W[r][c]
                W[r] = [a - f * b for a, b in zip(W[r], W[c])]
    return [row[n:] for row in W]


def solve_case(M, known, tau=None, reg=None):
    """One system.  Returns (acc as doubles, cond_inf as float)."""
    nc = len(known)
    A = [[F(0)] * nc for _ in range(nc)]
    for i in range(nc):
        for k in range(i + 1):
            # the sum M + reg is EVALUATED in double by the reference before LLT sees it (:236-239)
            v = F(float(M[i][k]) + float(reg[i][k])) if reg is not None else F(float(M[i][k]))
            A[i][k] = A[k][i] = v
    rhs = [F(float(x)) for x in known]
    if tau is not None:
        for q in range(6, nc):
            rhs[q] = F(float(known[q]) + float(tau[q - 6]))   # one double addition (:226-227)
    eye = [[F(int(i == k)) for k in range(nc)] for i in range(nc)]
    sol = _solve_exact(A, [[rhs[i]] + eye[i] for i in range(nc)])
    x = [row[0] for row in sol]
    inv = [row[1:] for row in sol]
    norm = lambda Mx: max(sum(abs(v) for v in row) for row in Mx)
    return np.array([float(v) for v in x]), float(norm(A) * norm(inv))


def main():
    from bipedal_locomotion_framework_b200 import synthetic as syn
    out = {}
    cases = [("a", 6, 24, 0.0, False, False), ("b", 12, 16, 0.5, True, False),
             ("c", 29, 8, 1.0, True, True), ("d", 18, 8, 1.5, False, True),
             ("e", 31, 4, 0.0, True, True), ("f", 40, 3, 0.5, True, True), ("g", 7, 16, 2.0, True, False)]
    for tag, nc, ns, spread, with_tau, with_reg in cases:
        rng = np.random.default_rng(1000 + nc)
        M = syn.make_mass_matrices(ns, nc, seed=300 + nc, spread=spread)
        known = rng.normal(size=(ns, nc)) * 50.0
        tau = rng.normal(size=(ns, nc - 6)) * 10.0 if (with_tau and nc > 6) else None
        reg = None
        if with_reg:
            reg = np.diag(10.0 ** rng.uniform(-4, -2, nc))
            reg[nc - 1, 0] = reg[0, nc - 1] = 1e-3          # not only diagonal
        acc = np.empty((ns, nc))
        cond = np.empty(ns)
        for s in range(ns):
            acc[s], cond[s] = solve_case(M[s], known[s], None if tau is None else tau[s], reg)
        out[tag + "_M"], out[tag + "_known"], out[tag + "_acc"], out[tag + "_cond"] = M, known, acc, cond
        if tau is not None:
            out[tag + "_tau"] = tau
        if reg is not None:
            out[tag + "_reg"] = reg
        print(tag, nc, ns, "cond_inf max %.3g" % cond.max())
    out["tags"] = np.array([c[0] for c in cases])
    path = os.path.join(_ROOT, "tests", "golden", "dyn_exact_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

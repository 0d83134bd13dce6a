"""Parity metric of SURVEY.md section 8(d): per state and per 3-vector block,
||got - ref||_inf / max(||ref||_inf, floor) <= tol, structural zeros of the control matrix exactly
+0.0.  north_star tolerance: 1e-12 relative in FP64."""
import numpy as np

TOL = 1e-12  # north_star: "within 1e-12 relative in FP64"

BLOCKS = {
    "wrench": [slice(0, 3), slice(3, 6)],                      # force, torque
    "autodyn": [slice(0, 3), slice(3, 6)],                     # f.head, f.tail
    # each row of the two 3x3 diagonal blocks of g
    "ctrl": [slice(6 * r + 3 * (r // 3), 6 * r + 3 * (r // 3) + 3) for r in range(6)],
    # the four 3x1 corners of the row-major 6x2 regressor
    "regressor": [slice(0, 6, 2), slice(1, 6, 2), slice(6, 12, 2), slice(7, 12, 2)],
}

CTRL_STRUCTURAL_ZERO = np.ones(36, dtype=bool)
for _i in range(3):
    CTRL_STRUCTURAL_ZERO[6 * _i + _i] = False
    for _j in range(3):
        CTRL_STRUCTURAL_ZERO[6 * (3 + _i) + 3 + _j] = False


def block_rel_err(got, ref, key, floor=1e-300):
    """(n,) worst block-wise relative error of each state.  floor may be scalar or (n,)."""
    got = np.asarray(got).reshape(ref.shape)
    worst = np.zeros(ref.shape[0])
    for s in BLOCKS[key]:
        num = np.abs(got[:, s] - ref[:, s]).max(axis=1)
        den = np.maximum(np.abs(ref[:, s]).max(axis=1), floor)
        worst = np.maximum(worst, np.where(num == 0.0, 0.0, num / den))
    return worst


def assert_parity(got, ref, key, tol=TOL, floor=1e-300, what=""):
    assert np.all(np.isfinite(got)), f"{what}{key}: non-finite output"
    err = block_rel_err(got, ref, key, floor)
    i = int(np.argmax(err))
    assert err[i] <= tol, f"{what}{key}: state {i} rel err {err[i]:.3e} > {tol:g}"
    return float(err.max()), int((err > 1e-13).sum())


def assert_ctrl_structure(ctrl):
    """24 structural zeros exactly +0.0 (never written after the constructor's zero(),
    ContinuousContactModel.cpp:18,165-170); the three top-left diagonal entries equal."""
    c = np.asarray(ctrl).reshape(-1, 36)
    z = c[:, CTRL_STRUCTURAL_ZERO]
    assert np.all(z == 0.0) and not np.signbit(z).any(), "structural zero is not +0.0"
    assert np.array_equal(c[:, 0], c[:, 7]) and np.array_equal(c[:, 0], c[:, 14])


EPS = 2.0 ** -52


def llt_tolerance(nc, cond):
    """Tolerance for a Cholesky solve of an nc x nc system with condition number `cond`: the flat
    north_star 1e-12 where the conditioning allows it, else the forward-error bound of any
    floating-point LLT, c n eps cond(A) (Higham, Accuracy and Stability of Numerical Algorithms,
    thm 10.4 + 7.2) -- the reference's own Eigen LLT is only that close to the exact solution."""
    return np.maximum(TOL, 4.0 * nc * EPS * np.asarray(cond))

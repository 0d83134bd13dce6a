"""CPU tests of the System-component oracle (oracle/sys_oracle.c): against the exact / 80-digit
golden fixtures and against the reference test's own property
(src/System/tests/IntegratorTest.cpp:80-126).  No GPU."""
import os

import numpy as np
import pytest

from oracle import sys_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-12


@pytest.fixture(scope="module")
def g():
    from oracle import ccm_oracle
    ccm_oracle.build()
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "sys_exact_golden.npz")))


def rel(got, ref, axis=-1, floor=1e-300):
    num = np.abs(np.asarray(got) - ref).max(axis=axis)
    den = np.maximum(np.abs(ref).max(axis=axis), floor)
    return np.where(num == 0, 0.0, num / den)


def test_single_steps_match_exact_rationals(g):
    n = g["step_twists"].shape[0]
    for i in range(n):
        pd, rd = so.kinematics_dynamics(g["step_rho"][i], g["step_twists"][i], g["step_rot"][i])
        assert np.array_equal(pd, g["step_pos_dot"][i])          # a copy of the twist: exact
        # rotation rate: norm-wise over the matrix, against the operands of its cancelling sums
        # (w x col, and rho/2 * ((R R^T)^-1 - I) whose bracket is a difference of O(1) numbers)
        scale = max(np.abs(g["step_rot_dot"][i]).max(), np.abs(g["step_twists"][i][3:]).max(),
                    g["step_rho"][i] / 2)
        assert np.abs(rd.reshape(9) - g["step_rot_dot"][i]).max() <= TOL * max(scale, 1e-300), i
        p, r = so.forward_euler_step(g["step_rho"][i], g["step_dT"][i], g["step_twists"][i],
                                     g["step_pos"][i], g["step_rot"][i])
        assert rel(p, g["step_pos_new"][i]) <= TOL
        assert rel(r.reshape(9), g["step_rot_new"][i]) <= TOL


def test_batch_step_equals_per_instance(g):
    tw = np.ascontiguousarray(g["step_twists"].T)
    p0 = np.ascontiguousarray(g["step_pos"].T)
    r0 = np.ascontiguousarray(g["step_rot"].T)
    p, r = so.euler_step_batch_soa(2.5, 1e-3, tw, p0, r0, nthreads=3)
    for i in (0, 5, 63):
        pi, ri = so.forward_euler_step(2.5, 1e-3, g["step_twists"][i], g["step_pos"][i],
                                       g["step_rot"][i])
        assert np.array_equal(p[:, i], pi) and np.array_equal(r[:, i], ri.reshape(9))


def test_reference_integrator_property():
    """IntegratorTest.cpp:80-126: identity start, constant twist, dT = 1e-4, 2 s; rotation follows
    AngleAxis(|w| t, w/|w|), position and joints are linear in t; tolerance 1e-3 (isApprox)."""
    rng = np.random.default_rng(7)
    twist = rng.uniform(-1, 1, 6)
    jv = rng.uniform(-1, 1, 20)
    dT, T = 1e-4, 2.0
    p, R, jp = np.zeros(3), np.eye(3), np.zeros(20)
    w = twist[3:]
    th = np.linalg.norm(w)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    steps = int(T / dT)
    for i in range(steps):
        if i % 500 == 0:
            t = dT * i
            Rex = np.eye(3) + np.sin(th * t) * K + (1 - np.cos(th * t)) * K @ K
            assert np.linalg.norm(R - Rex) <= 1e-3 * min(np.linalg.norm(R), np.linalg.norm(Rex))
            assert np.linalg.norm(p - t * twist[:3]) <= 1e-3 * max(np.linalg.norm(p), 1e-12) + 1e-15
            assert np.linalg.norm(jp - t * jv) <= 1e-3 * max(np.linalg.norm(jp), 1e-12) + 1e-15
        # integrator.integrate(0, dT) with sampling time dT: one step of dT
        n, p, R, jp = so.integrate(0.0, dT, 0.0, dT, twist, p, R, jv, jp)
        assert n == 1


def test_integrate_schedule_quirk():
    """FixedStepIntegrator.tpp:48-64: iterations = ceil((tf-t0)/dT); currentTime only advances
    inside the loop, so with >= 2 iterations the last step is tf - (t0 + dT*(iterations-2))."""
    assert np.allclose(so.integrate_schedule(0.1, 0.0, 0.1), [0.1])
    s = so.integrate_schedule(0.25, 0.0, 1.0)          # 4 iterations: 3 loop steps + last
    assert len(s) == 4 and np.array_equal(s[:3], [0.25] * 3) and s[3] == 1.0 - 0.25 * 2
    s = so.integrate_schedule(0.3, 1.0, 2.0)           # ceil(3.33) = 4
    assert len(s) == 4 and s[3] == 2.0 - (1.0 + 0.3 * 2)
    assert so.integrate_schedule(0.1, 1.0, 0.5) is None      # tf < t0
    assert so.integrate_schedule(0.0, 0.0, 1.0) is None      # dT <= 0
    assert so.integrate_schedule(-0.1, 0.0, 1.0) is None
    assert so.integrate_schedule(0.1, 1.0, 1.0) is None      # reference would not terminate
    tw = np.array([0.1, 0.2, 0.3, 0.0, 0.0, 0.0])
    n, p, R = so.integrate(0.0, 0.25, 0.0, 1.0, tw, np.zeros(3), np.eye(3))
    assert n == 4 and np.allclose(p, tw[:3] * (0.75 + 0.5))   # integrates 1.25 s, as the reference


def test_baumgarte_restores_orthonormality():
    R = np.eye(3) * 1.05
    for _ in range(2000):
        _, R = so.forward_euler_step(20.0, 1e-3, np.zeros(6), np.zeros(3), R)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-6


def test_rho_zero_with_a_singular_rotation_is_nan_in_the_reference():
    """FloatingBaseSystemKinematics.cpp:62-66 multiplies rho/2 into ((R R^T)^-1 - I) R even when rho
    is 0: a singular R makes the inverse inf/NaN and 0 * inf = NaN poisons the rotation rate.  The
    oracle (and the reference build, tests/test_reference_build.py) reproduce that; the CUDA backend
    deliberately does not (include/blf_ccm.h, blf_sys_kinematics_*: with rho == 0 the inverse is
    skipped) -- tests/test_gpu_sys.py pins the backend's side of this documented deviation.  For a
    regular R, rho = 0 adds exact zeros and both agree."""
    twist = np.array([0.1, -0.2, 0.3, 0.4, 0.5, -0.6])
    singular = np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 6.0], [0.5, -1.0, 0.25]])   # rank 2
    pd, rd = so.kinematics_dynamics(0.0, twist, singular.reshape(9))
    assert np.array_equal(pd, twist[:3]) and np.isnan(rd).all()
    regular = np.array([[0.9, -0.1, 0.2], [0.1, 1.1, 0.0], [-0.2, 0.05, 0.95]])
    pd, rd = so.kinematics_dynamics(0.0, twist, regular.reshape(9))
    want = -np.cross(regular.T, twist[3:]).T           # -R.colwise().cross(w)
    assert np.isfinite(rd).all() and np.abs(rd.reshape(3, 3) - want).max() <= 1e-15


def test_rollout_matches_80_digit_golden(g):
    nr, feet, H = (int(x) for x in g["ro_shape"])
    for nthreads in (1, 4):
        out = so.rollout(nr, feet, H, float(g["ro_dT"]), float(g["ro_rho"]), g["ro_twists"],
                         g["ro_pos0"], g["ro_rot0"], g["ro_null"], param_planes=g["ro_params"],
                         mask=7, wrench_ref=g["ro_ref"], weights=g["ro_weights"],
                         nthreads=nthreads)
        assert rel(out["pos"].T, g["ro_pos"].T).max() <= TOL
        assert rel(out["rot"].T, g["ro_rot"].T).max() <= TOL
        for key in ("wrench", "autodyn"):
            got, ref = out[key].T, g["ro_" + key].T
            for blk in (slice(0, 3), slice(3, 6)):
                assert rel(got[:, blk], ref[:, blk]).max() <= TOL, key
        c, cr = out["ctrl"], g["ro_ctrl"]
        assert rel(c[:, [0, 7, 14]], cr[:, [0, 7, 14]]).max() <= TOL
        idx = [21, 22, 23, 27, 28, 29, 33, 34, 35]
        assert rel(c[:, idx], cr[:, idx]).max() <= TOL
        assert rel(out["chain_cost"], g["ro_chain_cost"], axis=None) <= TOL
        assert rel(out["cost"], g["ro_cost"], axis=None) <= TOL


def test_rollout_uniform_and_cost_only(g):
    nr, feet, H = (int(x) for x in g["ro_shape"])
    uni = (0.12, 0.09, 2000.0, 100.0)
    full = so.rollout(nr, feet, H, 0.01, 0.0, g["ro_twists"], g["ro_pos0"], g["ro_rot0"],
                      g["ro_null"], uniform=uni, mask=1, wrench_ref=g["ro_ref"],
                      weights=g["ro_weights"])
    only = so.rollout(nr, feet, H, 0.01, 0.0, g["ro_twists"], g["ro_pos0"], g["ro_rot0"],
                      g["ro_null"], uniform=uni, mask=0, wrench_ref=g["ro_ref"],
                      weights=g["ro_weights"])
    assert np.array_equal(full["cost"], only["cost"]) and only["wrench"] is None
    # cost re-derived from the wrench trajectory
    chains = nr * feet
    w = full["wrench"].T.reshape(H, chains, 6)
    d = w - g["ro_ref"]
    term = g["ro_weights"][0] * (d[..., :3] ** 2).sum(-1) + g["ro_weights"][1] * (d[..., 3:] ** 2).sum(-1)
    assert np.allclose(term.sum(0).reshape(nr, feet).sum(1), full["cost"], rtol=1e-13)


@pytest.mark.parametrize("tag", ["gfa", "gfb"])
def test_generalized_force_matches_exact(g, tag):
    from bipedal_locomotion_framework_b200 import synthetic as syn
    ns, cps, ncols = (int(x) for x in g[tag + "_shape"])
    planes = syn.aos_to_planes(g[tag + "_twists"], g[tag + "_poses"], g[tag + "_null_poses"])
    prm = np.ascontiguousarray(g[tag + "_params"].T)
    out, wr = so.generalized_force(cps, ncols, planes, g[tag + "_J"], g[tag + "_base"],
                                   param_planes=prm, want_wrench=True, nthreads=2)
    # norm-wise per system against the magnitude of the summed terms
    J = g[tag + "_J"].reshape(ns, cps, 6, ncols)
    W = g[tag + "_wrench"].reshape(ns, cps, 6)
    mag = np.abs(g[tag + "_base"]) + np.einsum("scrq,scr->sq", np.abs(J), np.abs(W))
    assert (np.abs(out - g[tag + "_out"]).max(axis=1) <= TOL * mag.max(axis=1)).all()
    assert rel(wr.T[:, :3], g[tag + "_wrench"][:, :3]).max() <= TOL
    assert rel(wr.T[:, 3:], g[tag + "_wrench"][:, 3:]).max() <= TOL
    # base = NULL means zeros
    out0 = so.generalized_force(cps, ncols, planes, g[tag + "_J"], None, param_planes=prm)
    assert np.allclose(out0 + g[tag + "_base"], g[tag + "_out"], rtol=0, atol=1e-9 * mag.max())


# --- (M + reg).llt().solve(known + torques): the end of FloatingBaseDynamicalSystem::dynamics -------

EPS = 2.0 ** -52


def llt_tolerance(nc, cond):
    """Forward-error bound of a floating-point Cholesky solve: relative error <= c n eps cond(A)
    (Higham, Accuracy and Stability of Numerical Algorithms, thm 10.4 + 7.2); flat 1e-12 (north_star)
    where the conditioning allows it."""
    return np.maximum(TOL, 4.0 * nc * EPS * cond)


@pytest.fixture(scope="module")
def gd():
    from oracle import ccm_oracle
    ccm_oracle.build()
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "dyn_exact_golden.npz")))


def test_mass_matrix_solve_matches_exact_rationals(gd):
    """oracle/exact_golden_dyn.py: the solution is rational in the inputs, evaluated exactly."""
    for tag in gd["tags"]:
        M, known, acc, cond = (gd[f"{tag}_{k}"] for k in ("M", "known", "acc", "cond"))
        tau, reg = gd.get(f"{tag}_tau"), gd.get(f"{tag}_reg")
        nc = known.shape[1]
        for nthreads in (1, 3):
            x = so.mass_matrix_solve(M, known, tau, reg, nthreads=nthreads)
            assert (rel(x, acc) <= llt_tolerance(nc, cond)).all(), tag
        # well conditioned systems meet the flat tolerance
        easy = cond < 100
        if easy.any():
            assert rel(x[easy], acc[easy]).max() <= TOL


def test_mass_matrix_solve_reads_only_the_lower_triangle(gd):
    """Eigen's LLT<Lower>: the strict upper triangle never enters."""
    M, known = gd["c_M"].copy(), gd["c_known"]
    x0 = so.mass_matrix_solve(M, known)
    iu = np.triu_indices(M.shape[1], 1)
    M[:, iu[0], iu[1]] = np.nan
    assert np.array_equal(so.mass_matrix_solve(M, known), x0)


def test_mass_matrix_solve_not_positive_definite_is_nan():
    M = np.array([[[1.0, 2.0], [2.0, 1.0]]])
    assert np.isnan(so.mass_matrix_solve(M, np.ones((1, 2)))).any()


def test_floating_base_acceleration_composes_force_and_solve(g):
    from bipedal_locomotion_framework_b200 import synthetic as syn
    tag = "gfa"
    ns, cps, ncols = (int(x) for x in g[tag + "_shape"])
    planes = syn.aos_to_planes(g[tag + "_twists"], g[tag + "_poses"], g[tag + "_null_poses"])
    prm = np.ascontiguousarray(g[tag + "_params"].T)
    rng = np.random.default_rng(5)
    bias = rng.normal(size=(ns, ncols))
    tau = rng.normal(size=(ns, ncols - 6))
    M = syn.make_mass_matrices(ns, ncols, seed=2)
    acc = so.floating_base_acceleration(cps, planes, g[tag + "_J"], bias, M, tau, param_planes=prm)
    known = so.generalized_force(cps, ncols, planes, g[tag + "_J"], -bias, param_planes=prm)
    assert np.array_equal(acc, so.mass_matrix_solve(M, known, tau))
    # residual of the linear system, the property that does not depend on any factorisation
    rhs = known.copy()
    rhs[:, 6:] += tau
    res = np.einsum("sij,sj->si", M, acc) - rhs
    assert np.abs(res).max() <= 1e-12 * np.abs(rhs).max()


@pytest.mark.parametrize("nc", [1, 2, 6, 13, 29, 40])
def test_mass_matrix_solve_properties(nc):
    """Size-independent properties of the solve: M * solve(M, b) = b to rounding; a zero regularisation
    and zero torques change nothing, bit for bit; systems are independent (any order, any threads)."""
    from bipedal_locomotion_framework_b200 import synthetic as syn
    rng = np.random.default_rng(nc)
    ns = 64
    M = syn.make_mass_matrices(ns, nc, seed=nc, spread=0.5)
    xs = rng.normal(size=(ns, nc))
    b = np.einsum("sij,sj->si", M, xs)
    x = so.mass_matrix_solve(M, b)
    cond = np.linalg.cond(M, np.inf)
    assert (rel(x, xs) <= llt_tolerance(nc, cond)).all()
    tau0 = np.zeros((ns, nc - 6)) if nc > 6 else None
    assert np.array_equal(so.mass_matrix_solve(M, b, tau0, np.zeros((nc, nc))), x)
    perm = rng.permutation(ns)
    assert np.array_equal(so.mass_matrix_solve(M[perm], b[perm], nthreads=3), x[perm])
    # the regularisation enters as M + reg, evaluated once in double
    reg = np.diag(rng.uniform(1e-4, 1e-2, nc))
    assert np.array_equal(so.mass_matrix_solve(M, b, None, reg), so.mass_matrix_solve(M + reg, b))


@pytest.mark.parametrize("nc,spread", [(6, 0.0), (12, 1.0), (29, 0.5), (29, 1.5), (38, 1.0), (64, 0.5)])
def test_mass_matrix_solve_against_lapack(nc, spread):
    """An implementation nobody here wrote: numpy's LAPACK solve (dgesv, partial pivoting) on the same
    systems.  Eigen's LLT is absent from this image; LAPACK stands for 'any correct solver in IEEE
    double' and bounds what a real-Eigen build could differ by: rounding x cond(M)."""
    from bipedal_locomotion_framework_b200 import synthetic as syn
    rng = np.random.default_rng(100 + nc)
    ns = 500
    M = syn.make_mass_matrices(ns, nc, seed=9 + nc, spread=spread)
    b = rng.normal(size=(ns, nc)) * 40.0
    tau = rng.normal(size=(ns, nc - 6)) if nc > 6 else None
    rhs = b.copy()
    if tau is not None:
        rhs[:, 6:] += tau
    want = np.linalg.solve(M, rhs[:, :, None])[:, :, 0]
    x = so.mass_matrix_solve(M, b, tau, nthreads=2)
    cond = np.linalg.cond(M, np.inf)
    assert (rel(x, want) <= 2 * llt_tolerance(nc, cond)).all()
    if spread == 0.0:
        assert rel(x, want).max() <= TOL


@pytest.mark.parametrize("nc,rho", [(6, 0.0), (12, 0.3), (29, 0.7)])
def test_floating_base_euler_step_properties(nc, rho):
    """syso_floating_base_euler_step (ForwardEuler.tpp:19-49 over the floating-base state): the pose and
    joint part IS the kinematics' forward Euler step with the OLD velocity as control input, the velocity
    part is nu + acc dT with one rounding each, dT = 0 changes nothing, and the inputs are not written."""
    rng = np.random.default_rng(40 + nc)
    ns, dT = 37, 0.01
    acc, nu = rng.normal(size=(ns, nc)) * 20.0, rng.normal(size=(ns, nc))
    jp = rng.normal(size=(ns, nc - 6)) if nc > 6 else None
    p = rng.normal(size=(ns, 3))
    R = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] * (1.0 + 0.03 * rng.normal()) for _ in range(ns)])
    keep = [x.copy() if x is not None else None for x in (acc, nu, jp, p, R)]
    v, q, pp, RR = so.floating_base_euler_step(rho, dT, acc, nu, jp, p, R, nthreads=2)
    for a, b in zip((acc, nu, jp, p, R), keep):
        assert a is None or np.array_equal(a, b)
    assert np.array_equal(v, nu + acc * dT)
    for s in range(ns):
        out = so.forward_euler_step(rho, dT, nu[s, :6], p[s], R[s], nu[s, 6:] if nc > 6 else None,
                                    jp[s] if nc > 6 else None)
        kp, kR, kq = out if nc > 6 else (*out, None)
        assert np.array_equal(pp[s], kp) and np.array_equal(RR[s], np.asarray(kR).reshape(3, 3))
        if nc > 6:
            assert np.array_equal(q[s], kq)
    v0, q0, p0, R0 = so.floating_base_euler_step(rho, 0.0, acc, nu, jp, p, R)
    assert np.array_equal(v0, nu) and np.array_equal(p0, p) and np.array_equal(R0, R)
    assert nc == 6 or np.array_equal(q0, jp)

"""Estimators::RecursiveLeastSquare: oracle pinned on CPU, CUDA path
checked against it on the GPU.  Reference: src/Estimators/src/RecursiveLeastSquare.cpp:96-133,
test src/Estimators/tests/RecursiveLeastSquareTest.cpp:91-142.
Tolerance 1e-12 relative, norm-wise per quantity and measured against the operands of the final
update: theta_new = theta_old + K innov is compared relative to max(|theta_old|, |theta_new|), and
P_new = (P_old - K Y P_old) / lambda relative to |P_old| -- the reference's own expression cancels,
so any two faithful evaluations differ by eps * |P_old|, not eps * |P_new|."""
import os

import numpy as np
import pytest

from bipedal_locomotion_framework_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-12


def _rel(got, ref, old=None):
    got, ref = np.asarray(got), np.asarray(ref)
    n = ref.shape[0]
    num = np.abs(got - ref).reshape(n, -1).max(axis=1)
    den = np.abs(ref).reshape(n, -1).max(axis=1)
    if old is not None:
        den = np.maximum(den, np.abs(np.asarray(old)).reshape(n, -1).max(axis=1))
    return float((num / np.maximum(den, 1e-300)).max())


def _rel_each(got, ref, old):
    got, ref, old = np.asarray(got), np.asarray(ref), np.asarray(old)
    n = ref.shape[0]
    num = np.abs(got - ref).reshape(n, -1).max(axis=1)
    den = np.maximum(np.abs(ref).reshape(n, -1).max(axis=1), np.abs(old).reshape(n, -1).max(axis=1))
    return num / np.maximum(den, 1e-300)


def _tolerance(Y, P, r, lam):
    """Per-estimator bound: two backward-stable evaluations of K = P Y^T S^-1 (the reference's LU on
    S = lambda R + Y P Y^T, ours on I + P Y^T W Y) differ by O(eps * cond(S)); 1e-12 whenever
    cond(S) < ~70 (the constant covers two O(m^2) solves, one of them through an explicit inverse)."""
    S = lam * np.diag(r)[None] + Y @ P @ Y.transpose(0, 2, 1)
    return np.maximum(TOL, 128 * np.finfo(np.float64).eps * np.linalg.cond(S))


def _assert_step(got_theta, got_P, ref_theta, ref_P, old_theta, old_P, Y, z, r, lam, what):
    """theta_new = theta_old + (P Y^T) S^-1 innov is a chain of products whose rounding error scales
    with max(|theta_old|, |P Y^T| |S^-1| |innov|) (entry-wise absolute values), so that -- not
    |theta_new| alone -- is the denominator."""
    tol = _tolerance(Y, old_P, r, lam)
    n = Y.shape[0]
    S = lam * np.diag(r)[None] + Y @ old_P @ Y.transpose(0, 2, 1)
    PYt = np.abs(old_P @ Y.transpose(0, 2, 1))
    innov = z - np.einsum("nmp,np->nm", Y, old_theta)
    terms = np.einsum("npm,nmk,nk->np", PYt, np.abs(np.linalg.inv(S)), np.abs(innov)).max(axis=1)
    num_t = np.abs(got_theta - ref_theta).reshape(n, -1).max(axis=1)
    den_t = np.maximum.reduce([np.abs(ref_theta).max(axis=1), np.abs(old_theta).max(axis=1), terms])
    et = num_t / np.maximum(den_t, 1e-300)
    eP = _rel_each(got_P, ref_P, old_P)
    i, j = int(np.argmax(et / tol)), int(np.argmax(eP / tol))
    assert et[i] <= tol[i], f"{what}: theta rel err {et[i]:.3e} > {tol[i]:.3e} (estimator {i})"
    assert eP[j] <= tol[j], f"{what}: P rel err {eP[j]:.3e} > {tol[j]:.3e} (estimator {j})"
    return float(max(np.median(et), np.median(eP)))


@pytest.fixture(scope="module")
def rls_golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "rls_exact_golden.npz")))


def _reference_test_model(rng):
    """y = [[x, x^2], [sin x, cos x]] p + noise, RecursiveLeastSquareTest.cpp:35-89."""
    params = np.array([43.2, 12.2])
    x = [0.0]
    reg = lambda: np.array([[x[0], x[0] ** 2], [np.sin(x[0]), np.cos(x[0])]])
    out = lambda: reg() @ params + rng.normal(0.0, 0.5, 2)
    return params, x, reg, out


# --- CPU: the oracle is pinned -----------------------------------------------------------------

def test_rls_oracle_matches_exact_golden(oracle, rls_golden):
    g = rls_golden
    steps = g["z"].shape[1]
    for t in range(steps):  # every step from the golden's own (exactly rounded) previous state
        th, P = oracle.rls_advance_batch(g["Y"][:, t], g["z"][:, t], g["r"], float(g["lam"]),
                                         g["theta"][:, t], g["P"][:, t])
        assert _rel(th, g["theta"][:, t + 1], g["theta"][:, t]) <= 1e-13
        assert _rel(P, g["P"][:, t + 1], g["P"][:, t]) <= 1e-13


def test_rls_oracle_reference_convergence_property(oracle):
    """10 000 steps recover (43.2, 12.2) within 0.1 % (RecursiveLeastSquareTest.cpp:122-141);
    configuration of src/Estimators/tests/config.ini."""
    rng = np.random.default_rng(42)
    params, x, reg, out = _reference_test_model(rng)
    est = oracle.RecursiveLeastSquare()
    cfg = {"lambda": 1.0, "measurement_covariance": [0.5, 0.5], "state": [0.0, 0.0],
           "state_covariance": [10.0, 10.0]}
    assert est.initialize(cfg)
    assert not est.initialize(cfg)            # "already initialized"
    est.setRegressorFunction(reg)
    for i in range(10000):
        x[0] = np.cos(i / 10.0)
        est.setMeasurements(out())
        assert est.advance()
    assert np.all(np.abs(est.parametersExpectedValue() - params) / params < 1e-3)


# --- GPU -----------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


@pytest.fixture(scope="module")
def batch(torch):
    from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch
    b = ContinuousContactModelBatch(0)
    b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
    return b


def _handler(**kw):
    from bipedal_locomotion_framework_b200.contact_models import StdImplementation
    h = StdImplementation()
    for k, v in kw.items():
        h.setParameter(k if k != "lam" else "lambda", v)
    return h


@pytest.mark.gpu
def test_reference_rls_test_on_the_gpu_facade(torch):
    from bipedal_locomotion_framework_b200.estimators import RecursiveLeastSquare
    rng = np.random.default_rng(42)
    params, x, reg, out = _reference_test_model(rng)
    est = RecursiveLeastSquare()
    assert not est.advance()                                       # not initialised, no regressor
    assert not est.initialize(_handler(lam=1.0, state=[0.0, 0.0], state_covariance=[10.0, 10.0]))
    h = _handler(lam=1.0, measurement_covariance=[0.5, 0.5], state=[0.0, 0.0],
                 state_covariance=[10.0, 10.0])
    assert est.initialize(h)
    assert not est.initialize(h)                                   # already initialised
    est.setRegressorFunction(reg)
    for i in range(10000):
        x[0] = np.cos(i / 10.0)
        est.setMeasurements(out())
        assert est.advance()
    assert np.all(np.abs(est.parametersExpectedValue() - params) / params < 1e-3)
    P = est.parametersCovarianceMatrix()
    assert P.shape == (2, 2) and np.all(np.diag(P) > 0) and np.all(np.diag(P) < 10.0)


@pytest.mark.gpu
def test_rls_exact_golden_on_gpu(torch, batch, rls_golden):
    from bipedal_locomotion_framework_b200 import _capi
    from bipedal_locomotion_framework_b200.contact_models import _np_ptr
    g = rls_golden
    n, steps = g["z"].shape[:2]
    for t in range(steps):
        th = np.ascontiguousarray(g["theta"][:, t]).copy()
        P = np.ascontiguousarray(g["P"][:, t]).copy()
        Y = np.ascontiguousarray(g["Y"][:, t])
        z = np.ascontiguousarray(g["z"][:, t])
        r = np.ascontiguousarray(g["r"])
        _capi.check(_capi.lib().blf_rls_advance_host(batch.handle.ptr, n, 2, 6, _np_ptr(Y), _np_ptr(z),
                                                     _np_ptr(r), float(g["lam"]), _np_ptr(th),
                                                     _np_ptr(P)))
        assert _rel(th, g["theta"][:, t + 1], g["theta"][:, t]) <= TOL
        assert _rel(P, g["P"][:, t + 1], g["P"][:, t]) <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("p,m", [(1, 1), (2, 2), (2, 6), (3, 4), (4, 6), (1, 6), (4, 1)])
def test_rls_batch_sizes_against_oracle(torch, batch, oracle, p, m):
    from bipedal_locomotion_framework_b200.estimators import RecursiveLeastSquareBatch
    rng = np.random.default_rng(100 * p + m)
    n = 4097
    r = rng.uniform(0.1, 1.0, m)
    lam = 0.97
    rls = RecursiveLeastSquareBatch(batch, r, lam)
    theta = rng.normal(0, 1, (n, p))
    A = rng.normal(0, 1, (n, p, p))
    P = A @ A.transpose(0, 2, 1) + 0.5 * np.eye(p)                 # SPD covariances
    d_theta = torch.from_numpy(np.ascontiguousarray(theta.T)).cuda()
    d_P = torch.from_numpy(np.ascontiguousarray(P.reshape(n, p * p).T)).cuda()
    for step in range(5):
        Y = rng.normal(0, 1, (n, m, p))
        z = rng.normal(0, 1, (n, m))
        theta_old, P_old = theta, P
        theta, P = oracle.rls_advance_batch(Y, z, r, lam, theta, P)
        rls.advance(torch.from_numpy(np.ascontiguousarray(Y.reshape(n, m * p).T)).cuda(),
                    torch.from_numpy(np.ascontiguousarray(z.T)).cuda(), d_theta, d_P)
        med = _assert_step(d_theta.cpu().numpy().T, d_P.cpu().numpy().T.reshape(n, p, p), theta, P,
                           theta_old, P_old, Y, z, r, lam, f"p={p} m={m} step {step}")
        assert med <= 1e-13      # the typical estimator agrees far below the bound
        # continue both chains from identical bits
        d_theta.copy_(torch.from_numpy(np.ascontiguousarray(theta.T)))
        d_P.copy_(torch.from_numpy(np.ascontiguousarray(P.reshape(n, p * p).T)))


@pytest.mark.gpu
@pytest.mark.parametrize("p,m,negative", [(5, 3, False), (8, 8, False), (6, 7, False), (2, 8, False),
                                          (2, 6, True), (4, 4, True), (3, 1, True)])
@pytest.mark.parametrize("layout", ["soa", "host"])
def test_rls_general_sizes_and_indefinite_covariance_against_oracle(torch, batch, oracle, p, m, negative, layout):
    """What the reference accepts and the register kernels do not cover -- more than 4 parameters
    or 6 measurements, or a lambda*R that is not positive (S not positive definite) -- runs the
    reference's own algorithm (LU with partial pivoting) in rls_advance_generic_kernel."""
    from bipedal_locomotion_framework_b200.estimators import RecursiveLeastSquareBatch
    import ctypes as C
    from bipedal_locomotion_framework_b200 import _capi
    rng = np.random.default_rng(1000 * p + 10 * m + negative)
    n = 1025
    r = rng.uniform(0.1, 1.0, m)
    if negative:
        r[rng.integers(0, m)] *= -1.0           # an indefinite measurement covariance entry
    lam = 0.97
    theta = rng.normal(0, 1, (n, p))
    A = rng.normal(0, 1, (n, p, p))
    P = A @ A.transpose(0, 2, 1) + 0.5 * np.eye(p)
    for step in range(3):
        Y = rng.normal(0, 1, (n, m, p))
        z = rng.normal(0, 1, (n, m))
        theta_ref, P_ref = oracle.rls_advance_batch(Y, z, r, lam, theta, P)
        if layout == "soa":
            d_theta = torch.from_numpy(np.ascontiguousarray(theta.T)).cuda()
            d_P = torch.from_numpy(np.ascontiguousarray(P.reshape(n, p * p).T)).cuda()
            RecursiveLeastSquareBatch(batch, r, lam).advance(
                torch.from_numpy(np.ascontiguousarray(Y.reshape(n, m * p).T)).cuda(),
                torch.from_numpy(np.ascontiguousarray(z.T)).cuda(), d_theta, d_P)
            got_t, got_P = d_theta.cpu().numpy().T, d_P.cpu().numpy().T.reshape(n, p, p)
        else:
            got_t, got_P = theta.copy(), np.ascontiguousarray(P).copy()
            ptr = lambda a: a.ctypes.data_as(C.c_void_p)
            Yc, zc, rc_ = np.ascontiguousarray(Y), np.ascontiguousarray(z), np.ascontiguousarray(r)
            _capi.check(_capi.lib().blf_rls_advance_host(batch.handle.ptr, n, p, m, ptr(Yc), ptr(zc), ptr(rc_),
                                                         lam, ptr(got_t), ptr(got_P)))
        # same algorithm, same order of operations as the oracle: agreement far below the bound
        _assert_step(got_t, got_P, theta_ref, P_ref, theta, P, Y, z, r, lam, f"p={p} m={m} step {step} {layout}")
        theta, P = theta_ref, P_ref


@pytest.mark.gpu
def test_rls_sizes_beyond_the_oracle_against_numpy(torch, batch):
    """p = 12 parameters, m = 10 measurements through the per-instance facade (the reference takes
    dynamic sizes): compared with a numpy restatement of RecursiveLeastSquare.cpp:118-130."""
    from bipedal_locomotion_framework_b200.estimators import RecursiveLeastSquare
    rng = np.random.default_rng(5)
    p, m, lam = 12, 10, 0.99
    r = rng.uniform(0.2, 1.0, m)
    est = RecursiveLeastSquare()
    assert est.initialize(_handler(lam=lam, measurement_covariance=list(r), state=[0.0] * p,
                                   state_covariance=[4.0] * p))
    cur = {}
    est.setRegressorFunction(lambda: cur["Y"])
    theta, P = np.zeros(p), 4.0 * np.eye(p)
    truth = rng.normal(size=p)
    for _ in range(40):
        Y = rng.normal(size=(m, p))
        z = Y @ truth + 1e-3 * rng.normal(size=m)
        cur["Y"] = Y
        est.setMeasurements(z)
        assert est.advance()
        K = P @ Y.T @ np.linalg.inv(lam * np.diag(r) + Y @ P @ Y.T)
        theta = theta + K @ (z - Y @ theta)
        P = (P - K @ Y @ P) / lam
        assert np.abs(est.parametersExpectedValue() - theta).max() <= 1e-9 * max(1.0, np.abs(theta).max())
        assert np.abs(est.parametersCovarianceMatrix() - P).max() <= 1e-9 * np.abs(P).max()
    assert np.abs(est.parametersExpectedValue() - truth).max() < 1e-2


@pytest.mark.gpu
def test_fused_contact_identification(torch, batch, oracle):
    """blf_ccm_rls_advance_contacts == regressor kernel + blf_rls_advance_batch (bit for bit), and
    agrees with the oracle chain regressor -> advance; over many steps it recovers each contact's
    own (spring, damper) from noisy wrenches."""
    from bipedal_locomotion_framework_b200.estimators import RecursiveLeastSquareBatch
    n, steps = 20000, 60
    rng = np.random.default_rng(5)
    true_k = rng.uniform(1e3, 1e5, n)
    true_b = rng.uniform(10.0, 1e3, n)
    geom = np.stack([rng.uniform(0.08, 0.3, n), rng.uniform(0.04, 0.15, n)], axis=0)
    r = np.array([1.0, 1.0, 1.0, 1e-2, 1e-2, 1e-2])
    lam = 1.0
    rls = RecursiveLeastSquareBatch(batch, r, lam)
    mk = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    th_f, th_u = mk(np.stack([0.5 * true_k, 2.0 * true_b])), None
    P_f = mk(np.stack([np.full(n, 1e10), np.zeros(n), np.zeros(n), np.full(n, 1e6)]))
    th_u, P_u = th_f.clone(), P_f.clone()
    d_geom = mk(geom)
    for t in range(steps):
        st = syn.make_states(n, seed=900 + t)
        st["params"] = np.ascontiguousarray(np.stack([geom[0], geom[1], true_k, true_b], axis=1))
        ref = oracle.eval_batch_states(st, mask=1 | 8, nthreads=os.cpu_count() or 1)
        z = ref["wrench"] + rng.normal(0, 1.0, (n, 6)) * np.sqrt(r)
        planes = mk(syn.aos_to_planes(st["twists"], st["poses"], st["null_poses"]))
        d_z = mk(z.T)
        if t < 3:
            th_prev = th_f.cpu().numpy().T.copy()
            P_prev = P_f.cpu().numpy().T.reshape(n, 2, 2).copy()
        rls.advance_contacts(planes, d_z, th_f, P_f, geometry_planes=d_geom)
        if t < 3:
            # unfused: regressor through HBM, then the generic batch update
            prm = mk(np.stack([geom[0], geom[1], np.zeros(n), np.zeros(n)]))
            out = batch.evaluate_soa(planes, prm, 8)
            rls.advance(out["regressor"], d_z, th_u, P_u)
            assert torch.equal(th_f, th_u) and torch.equal(P_f, P_u)
            # oracle step from the SAME bits the GPU started this step from
            Yo = ref["regressor"].reshape(n, 6, 2)
            th_o, P_o = oracle.rls_advance_batch(Yo, z, r, lam, th_prev, P_prev)
            _assert_step(th_f.cpu().numpy().T, P_f.cpu().numpy().T.reshape(n, 2, 2), th_o, P_o,
                         th_prev, P_prev, Yo, z, r, lam, f"fused step {t}")
    est = th_f.cpu().numpy()
    assert np.median(np.abs(est[0] - true_k) / true_k) < 0.05
    assert np.median(np.abs(est[1] - true_b) / true_b) < 0.05


@pytest.mark.gpu
def test_pipelined_and_plain_kernels_and_prebound_calls_agree_bit_for_bit(torch, monkeypatch):
    """rls_advance_pipe_kernel (default, per-thread cp.async ring over a resident grid) and the plain
    one-estimator-per-thread kernel (BLF_CCM_TUNE_RLS_PIPE=1) share the arithmetic body: identical
    bits, at ragged sizes around the tile (128) and grid (3 blocks x SMs) boundaries; the pre-bound
    call forms launch the same thing."""
    from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch
    from bipedal_locomotion_framework_b200.estimators import RecursiveLeastSquareBatch
    rng = np.random.default_rng(77)
    r, lam = rng.uniform(0.1, 1.0, 6), 0.97

    def run(n, plain, prebound):
        if plain:
            monkeypatch.setenv("BLF_CCM_TUNE_RLS_PIPE", "1")
        else:
            monkeypatch.delenv("BLF_CCM_TUNE_RLS_PIPE", raising=False)
        b = ContinuousContactModelBatch(0)
        b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
        rls = RecursiveLeastSquareBatch(b, r, lam)
        g = np.random.default_rng(n)
        Y, z = g.normal(0, 1, (12, n)), g.normal(0, 1, (6, n))
        th = g.normal(0, 1, (2, n))
        A = g.normal(0, 1, (n, 2, 2))
        P = (A @ A.transpose(0, 2, 1) + 0.5 * np.eye(2)).reshape(n, 4).T
        d = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (Y, z, th, P)]
        for _ in range(3):
            if prebound:
                rls.prepare_advance(*d)()
            else:
                rls.advance(*d)
        torch.cuda.synchronize()
        return d[2].cpu().numpy(), d[3].cpu().numpy()

    sm = torch.cuda.get_device_properties(0).multi_processor_count
    for n in (1, 127, 128, 129, 128 * 3 * sm - 1, 128 * 3 * sm + 1, 128 * 3 * sm * 2 + 77):
        th_a, P_a = run(n, plain=False, prebound=False)
        th_b, P_b = run(n, plain=True, prebound=False)
        th_c, P_c = run(n, plain=False, prebound=True)
        assert np.array_equal(th_a, th_b) and np.array_equal(P_a, P_b), n
        assert np.array_equal(th_a, th_c) and np.array_equal(P_a, P_c), n

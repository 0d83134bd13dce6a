"""GPU parity tests for the steps either side of the contact model (SURVEY.md section 8(f) rows 2
and 3): FloatingBaseSystemKinematics + ForwardEuler, the fused integrate -> contact -> cost
rollout, and the J^T * wrench accumulation -- the CUDA path through the C ABI against the CPU
oracle on the same bits and against the exact / 80-digit golden fixtures.
Tolerance 1e-12 norm-wise relative (north_star)."""
import os

import numpy as np
import pytest

from bipedal_locomotion_framework_b200 import synthetic as syn
from parity import TOL, assert_ctrl_structure, assert_parity

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NTHREADS = max(1, (os.cpu_count() or 1))


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch


@pytest.fixture(scope="module")
def batch(torch):
    from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch
    b = ContinuousContactModelBatch(0)
    b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
    return b


@pytest.fixture(scope="module")
def so(oracle):
    from oracle import sys_oracle
    return sys_oracle


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "sys_exact_golden.npz")))


def _dev(torch, a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel(got, ref, axis=-1, floor=1e-300):
    num = np.abs(np.asarray(got) - ref).max(axis=axis)
    den = np.maximum(np.abs(ref).max(axis=axis), floor)
    return np.where(num == 0, 0.0, num / den)


# --- kinematics ------------------------------------------------------------------------------------

def test_euler_step_matches_exact_rationals(torch, batch, g):
    from bipedal_locomotion_framework_b200.system import KinematicsBatch
    kb = KinematicsBatch(0, batch.handle)
    n = g["step_twists"].shape[0]
    # rho and dT are per launch: run each golden state through its own launch group
    for i in range(n):
        tw = _dev(torch, g["step_twists"][i:i + 1].T)
        p = _dev(torch, g["step_pos"][i:i + 1].T)
        r = _dev(torch, g["step_rot"][i:i + 1].T)
        kb.euler_step(g["step_rho"][i], g["step_dT"][i], tw, p, r)
        assert rel(p.cpu().numpy()[:, 0], g["step_pos_new"][i]) <= TOL
        assert rel(r.cpu().numpy()[:, 0], g["step_rot_new"][i]) <= TOL


@pytest.mark.parametrize("rho", [0.0, 7.5])
def test_euler_step_batch_vs_oracle(torch, batch, so, rho):
    from bipedal_locomotion_framework_b200.system import KinematicsBatch
    kb = KinematicsBatch(0, batch.handle)
    n = 100_003                                        # ragged
    st = syn.make_states(n, seed=91)
    tw = np.ascontiguousarray(st["twists"].T)
    p0 = np.ascontiguousarray(st["poses"][:, :3].T)
    r0 = np.ascontiguousarray(st["poses"][:, 3:].T)
    p_ref, r_ref = so.euler_step_batch_soa(rho, 2e-3, tw, p0, r0, nthreads=NTHREADS)
    p, r = _dev(torch, p0), _dev(torch, r0)
    kb.euler_step(rho, 2e-3, _dev(torch, tw), p, r)
    assert rel(p.cpu().numpy().T, p_ref.T).max() <= TOL
    assert rel(r.cpu().numpy().T, r_ref.T).max() <= TOL


def test_rho_zero_skips_the_inverse_documented_deviation(torch, batch, so):
    """The one documented behavioural deviation of the System rows (include/blf_ccm.h,
    blf_sys_kinematics_*): with rho == 0 the backend does not form (R R^T)^-1 at all, so a SINGULAR
    rotation matrix gives the finite rate -R.colwise().cross(w) where the reference's
    0 * ((R R^T)^-1 - I) R is NaN (FloatingBaseSystemKinematics.cpp:62-66; pinned on the oracle side
    by tests/test_sys_oracle.py::test_rho_zero_with_a_singular_rotation_is_nan_in_the_reference).
    With rho != 0 a singular R is non-finite in both; for regular R both agree for every rho."""
    from bipedal_locomotion_framework_b200.system import KinematicsBatch
    kb = KinematicsBatch(0, batch.handle)
    twist = np.array([0.1, -0.2, 0.3, 0.4, 0.5, -0.6])
    singular = np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 6.0], [0.5, -1.0, 0.25]])
    regular = np.array([[0.9, -0.1, 0.2], [0.1, 1.1, 0.0], [-0.2, 0.05, 0.95]])
    R = np.stack([singular.reshape(9), regular.reshape(9)], axis=1)          # (9, 2) planes
    tw = np.repeat(twist[:, None], 2, axis=1)
    dT = 1e-2
    for rho in (0.0, 3.0):
        p, r = _dev(torch, np.zeros((3, 2))), _dev(torch, R)
        kb.euler_step(rho, dT, _dev(torch, tw), p, r)
        r = r.cpu().numpy()
        _, r_ref = so.euler_step_batch_soa(rho, dT, tw, np.zeros((3, 2)), R)
        assert rel(r[:, 1], r_ref[:, 1]) <= TOL                                  # regular: parity
        if rho == 0.0:
            assert np.isnan(r_ref[:, 0]).all()                                   # the reference's NaN
            want = singular + dT * (-np.cross(singular.T, twist[3:]).T)
            assert np.isfinite(r[:, 0]).all() and np.abs(r[:, 0] - want.reshape(9)).max() <= 1e-15
        else:
            assert not np.isfinite(r_ref[:, 0]).all() and not np.isfinite(r[:, 0]).all()


def test_per_instance_facade_reference_property(torch):
    """IntegratorTest.cpp:80-126 through the Python mirror of the facade (every step on the GPU):
    identity start, constant twist; 200 integrate(0, dT) calls, then the closed form."""
    from bipedal_locomotion_framework_b200.contact_models import StdImplementation
    from bipedal_locomotion_framework_b200.system import FloatingBaseSystemKinematics, ForwardEuler
    rng = np.random.default_rng(5)
    twist, jv = rng.uniform(-1, 1, 6), rng.uniform(-1, 1, 20)
    sysm = FloatingBaseSystemKinematics(0)
    assert sysm.initalize(None) is False
    h = StdImplementation()
    assert sysm.initalize(h) is False                      # "rho" missing
    h.setParameter("rho", 0.0)
    assert sysm.initalize(h)
    sysm.setControlInput((twist, jv))
    sysm.setState((np.zeros(3), np.eye(3), np.zeros(20)))
    ok, (pd, rd, sd) = sysm.dynamics()
    assert ok and np.array_equal(pd, twist[:3]) and np.array_equal(sd, jv)
    w = twist[3:]
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    assert np.allclose(rd, K, atol=1e-15)                  # Rdot = S(w) R at R = I
    integ = ForwardEuler(1e-4)
    assert integ.setDynamicalSystem(sysm) and not integ.setDynamicalSystem(sysm)
    for _ in range(200):
        assert integ.integrate(0, 1e-4)
    t = 200 * 1e-4
    p, R, s = integ.getSolution()
    th = np.linalg.norm(w)
    k = K / th
    Rex = np.eye(3) + np.sin(th * t) * k + (1 - np.cos(th * t)) * k @ k
    assert np.linalg.norm(R - Rex) <= 1e-3 * np.linalg.norm(Rex)
    assert np.allclose(p, t * twist[:3], rtol=1e-12) and np.allclose(s, t * jv, rtol=1e-12)
    # schedule quirk: integrate(0, 1.0) with dT 0.25 covers 1.25 s (FixedStepIntegrator.tpp:48-64)
    sysm.setState((np.zeros(3), np.eye(3), np.zeros(20)))
    integ2 = ForwardEuler(0.25)
    integ2.setDynamicalSystem(sysm)
    assert integ2.integrate(0.0, 1.0)
    assert np.allclose(integ2.getSolution()[0], 1.25 * twist[:3], rtol=1e-14)
    assert not integ2.integrate(1.0, 0.5) and not integ2.integrate(1.0, 1.0)


def test_per_instance_integrate_vs_oracle(torch, so):
    from bipedal_locomotion_framework_b200.contact_models import StdImplementation
    from bipedal_locomotion_framework_b200.system import FloatingBaseSystemKinematics, ForwardEuler
    rng = np.random.default_rng(6)
    st = syn.make_states(4, seed=77)
    for i in range(4):
        twist, jv = st["twists"][i], rng.uniform(-1, 1, 7)
        R0 = st["poses"][i, 3:].reshape(3, 3)
        p0, s0 = st["poses"][i, :3], rng.uniform(-1, 1, 7)
        rho = [0.0, 3.0, 20.0, 0.5][i]
        sysm = FloatingBaseSystemKinematics(0)
        h = StdImplementation()
        h.setParameter("rho", rho)
        assert sysm.initalize(h)
        sysm.setControlInput((twist, jv))
        sysm.setState((p0, R0, s0))
        integ = ForwardEuler(0.003)
        integ.setDynamicalSystem(sysm)
        assert integ.integrate(0.1, 0.2)
        n, p_ref, R_ref, s_ref = so.integrate(rho, 0.003, 0.1, 0.2, twist, p0, R0, jv, s0)
        assert n == 34
        p, R, s = integ.getSolution()
        assert rel(p, p_ref) <= TOL and rel(R.reshape(9), R_ref.reshape(9)) <= TOL
        assert rel(s, s_ref) <= TOL


# --- fused rollout ---------------------------------------------------------------------------------

def _check_traj(out, ref, mask, what):
    if mask & 1:
        assert_parity(out["wrench"].cpu().numpy().T, ref["wrench"].T, "wrench", what=what + " ")
    if mask & 2:
        assert_parity(out["autodyn"].cpu().numpy().T, ref["autodyn"].T, "autodyn", what=what + " ")
    if mask & 4:
        c = out["ctrl"].cpu().numpy()
        assert_parity(c, ref["ctrl"], "ctrl", what=what + " ")
        assert_ctrl_structure(c)


def test_rollout_matches_80_digit_golden(torch, batch, g):
    from bipedal_locomotion_framework_b200.system import RolloutBatch
    rb = RolloutBatch(batch)
    nr, feet, H = (int(x) for x in g["ro_shape"])
    out = rb.run(nr, feet, H, float(g["ro_dT"]), float(g["ro_rho"]), _dev(torch, g["ro_twists"]),
                 _dev(torch, g["ro_pos0"]), _dev(torch, g["ro_rot0"]), _dev(torch, g["ro_null"]),
                 g["ro_ref"], g["ro_weights"], param_planes=_dev(torch, g["ro_params"]), mask=7,
                 want_final=True)
    ref = {k[3:]: v for k, v in g.items() if k.startswith("ro_")}
    _check_traj(out, ref, 7, "rollout/golden")
    assert rel(out["final_pos"].cpu().numpy().T, ref["pos"].T).max() <= TOL
    assert rel(out["final_rot"].cpu().numpy().T, ref["rot"].T).max() <= TOL
    assert rel(out["cost"].cpu().numpy(), ref["cost"], axis=None) <= TOL
    cost, idx = batch.decode_best(out["best"])
    assert idx == int(np.argmin(ref["cost"])) and abs(cost - ref["cost"].min()) <= TOL * ref["cost"].min()


@pytest.mark.parametrize("het,rho,mask,nr,feet,H", [
    (False, 0.0, 0, 4096, 2, 100),     # configs[2] shape, cost only
    (False, 0.0, 7, 512, 2, 50),
    (True, 4.0, 7, 333, 3, 37),        # ragged: 999 chains, odd horizon, Baumgarte on
    (True, 4.0, 1, 1000, 1, 9),        # horizon shorter than... the prefetch ring + 1
    (False, 1.0, 5, 70, 2, 3),         # horizon < prefetch depth
])
def test_rollout_vs_oracle(torch, batch, so, het, rho, mask, nr, feet, H):
    from bipedal_locomotion_framework_b200.system import RolloutBatch
    rb = RolloutBatch(batch)
    chains = nr * feet
    st = syn.make_states(chains, seed=123, heterogeneous=het)
    tw = np.ascontiguousarray(syn.make_states(H * chains, seed=124)["twists"].T)
    pos0 = np.ascontiguousarray(st["poses"][:, :3].T)
    rot0 = np.ascontiguousarray(st["poses"][:, 3:].T)
    null = np.ascontiguousarray(st["null_poses"].T)
    prm = np.ascontiguousarray(st["params"].T) if het else None
    ref_w, wts = np.array([0.0, 0.0, 30.0, 0.1, -0.1, 0.0]), np.array([1.0, 25.0])
    dT = 0.01
    ref = so.rollout(nr, feet, H, dT, rho, tw, pos0, rot0, null, param_planes=prm,
                     uniform=syn.REFERENCE_TEST_PARAMS, mask=mask, wrench_ref=ref_w, weights=wts,
                     nthreads=NTHREADS)
    # the dead third column of the null rotation may be absent
    null_list = [None if i in (5, 8, 11) else _dev(torch, null[i]) for i in range(12)]
    out = rb.run(nr, feet, H, dT, rho, _dev(torch, tw), _dev(torch, pos0), _dev(torch, rot0),
                 null_list, ref_w, wts, param_planes=_dev(torch, prm), mask=mask, want_final=True)
    _check_traj(out, ref, mask, f"rollout het={het}")
    assert rel(out["final_pos"].cpu().numpy().T, ref["pos"].T).max() <= TOL
    assert rel(out["final_rot"].cpu().numpy().T, ref["rot"].T).max() <= TOL
    cost = out["cost"].cpu().numpy()
    assert np.all(rel(cost[:, None], ref["cost"][:, None]) <= TOL)
    c, idx = batch.decode_best(out["best"])
    assert c == cost.min() and idx == int(np.argmin(cost))
    # deterministic run to run (same kernel, same bits) ...
    out2 = rb.run(nr, feet, H, dT, rho, _dev(torch, tw), _dev(torch, pos0), _dev(torch, rot0),
                  null_list, ref_w, wts, param_planes=_dev(torch, prm), mask=mask)
    assert np.array_equal(out2["cost"].cpu().numpy(), cost)
    # ... and the cost-only variant (another template instance, other FMA contraction) agrees
    out3 = rb.run(nr, feet, H, dT, rho, _dev(torch, tw), _dev(torch, pos0), _dev(torch, rot0),
                  null_list, ref_w, wts, param_planes=_dev(torch, prm), mask=0)
    assert np.all(rel(out3["cost"].cpu().numpy()[:, None], ref["cost"][:, None]) <= TOL)


def test_rollout_equals_unfused_pipeline(torch, batch):
    """Size-independent property at configs[2] size: the fused rollout's wrench trajectory equals
    stepping KinematicsBatch + evaluate_soa one horizon step at a time (same arithmetic; the
    compiler may contract differently per kernel, so compared at the parity tolerance)."""
    from bipedal_locomotion_framework_b200.system import KinematicsBatch, RolloutBatch
    rb, kb = RolloutBatch(batch), KinematicsBatch(0, batch.handle)
    nr, feet, H = 4096, 2, 100
    chains = nr * feet
    st = syn.make_states(chains, seed=321)
    tw = _dev(torch, syn.make_states(H * chains, seed=322)["twists"].T)
    pos = _dev(torch, st["poses"][:, :3].T)
    rot = _dev(torch, st["poses"][:, 3:].T)
    null = _dev(torch, st["null_poses"].T)
    ref_w, wts = np.zeros(6), np.array([1.0, 1.0])
    out = rb.run(nr, feet, H, 0.005, 2.0, tw, pos, rot, null, ref_w, wts, mask=1, want_final=True)
    p, r = pos.clone(), rot.clone()
    for t in range(H):
        twt = tw[:, t * chains:(t + 1) * chains]
        planes = [twt[i] for i in range(6)] + [p[i] for i in range(3)] + [r[i] for i in range(9)] + \
                 [null[i] for i in range(12)]
        w = batch.evaluate_soa(planes, None, 1)["wrench"]
        assert_parity(out["wrench"][:, t * chains:(t + 1) * chains].cpu().numpy().T,
                      w.cpu().numpy().T, "wrench", what=f"step {t} ")
        kb.euler_step(2.0, 0.005, twt, p, r)
    assert rel(out["final_pos"].cpu().numpy().T, p.cpu().numpy().T).max() <= TOL
    assert rel(out["final_rot"].cpu().numpy().T, r.cpu().numpy().T).max() <= TOL


def test_rollout_argument_errors(torch, batch):
    from bipedal_locomotion_framework_b200 import _capi
    from bipedal_locomotion_framework_b200.system import RolloutBatch
    rb = RolloutBatch(batch)
    z = lambda *s: torch.zeros(s, dtype=torch.float64, device="cuda")
    with pytest.raises(_capi.BlfCcmError):
        rb.run(4, 0, 5, 0.01, 0.0, z(6, 20), z(3, 4), z(9, 4), z(12, 4), np.zeros(6), np.ones(2))
    with pytest.raises(_capi.BlfCcmError):
        rb.run(4, 1, 5, 0.01, 0.0, z(6, 20), z(3, 4), z(9, 4), z(12, 4), np.zeros(6), np.ones(2),
               mask=8)
    out = rb.run(0, 2, 5, 0.01, 0.0, z(6, 0), z(3, 0), z(9, 0), z(12, 0), np.zeros(6), np.ones(2))
    assert batch.decode_best(out["best"])[1] == -1


# --- J^T wrench ------------------------------------------------------------------------------------

@pytest.mark.parametrize("tag", ["gfa", "gfb"])
def test_generalized_force_matches_exact(torch, batch, g, tag):
    from bipedal_locomotion_framework_b200.system import GeneralizedForceBatch
    gf = GeneralizedForceBatch(batch)
    ns, cps, ncols = (int(x) for x in g[tag + "_shape"])
    planes = _dev(torch, syn.aos_to_planes(g[tag + "_twists"], g[tag + "_poses"],
                                           g[tag + "_null_poses"]))
    out, wr = gf.run(cps, ncols, planes, _dev(torch, g[tag + "_J"]), _dev(torch, g[tag + "_base"]),
                     param_planes=_dev(torch, g[tag + "_params"].T), want_wrench=True)
    J = g[tag + "_J"].reshape(ns, cps, 6, ncols)
    Wm = g[tag + "_wrench"].reshape(ns, cps, 6)
    mag = np.abs(g[tag + "_base"]) + np.einsum("scrq,scr->sq", np.abs(J), np.abs(Wm))
    assert (np.abs(out.cpu().numpy() - g[tag + "_out"]).max(axis=1) <= TOL * mag.max(axis=1)).all()
    assert_parity(wr.cpu().numpy().T, g[tag + "_wrench"], "wrench")


@pytest.mark.parametrize("cps,ncols,ns,het,aligned,with_base", [
    (2, 29, 50_001, False, True, True),     # iCub-sized: 6 + 23 DoF, two feet; ragged system count
    (2, 29, 4_099, True, False, True),      # Jacobians only 8-byte aligned: direct path
    (1, 6, 10_000, False, True, False),     # base-only Jacobian, no bias
    (3, 38, 3_000, True, True, True),       # 32 % 3 != 0, two column chunks
    (32, 128, 65, False, True, True),       # maxima
    (5, 1, 777, False, True, True),         # single column
    (2, 12, 20_001, True, True, True),      # narrow Jacobian: two contacts per warp iteration
    (3, 16, 5_000, False, False, True),     # ... on the direct (8-byte aligned) path
    (1, 9, 33_333, False, True, False),     # ... one contact per system
    (7, 8, 4_001, True, True, True),        # four contacts per iteration, 32 % 7 != 0
    (32, 5, 100, False, True, True),        # a whole warp is one system
    (8, 6, 4_001, False, True, True),       # 8 corner contacts of a rigid body: one 9 KB stage per warp
    (8, 8, 1_000, True, True, True),        # ... 12 KB, the largest single stage
    (16, 4, 999, False, True, False),       # two systems per warp on 4-lane groups
    (5, 6, 2_222, True, True, True),        # 6 systems per warp > 4 lane groups: two-stage ring
])
def test_generalized_force_vs_oracle(torch, batch, so, cps, ncols, ns, het, aligned, with_base):
    from bipedal_locomotion_framework_b200.system import GeneralizedForceBatch
    gf = GeneralizedForceBatch(batch)
    n = ns * cps
    st = syn.make_states(n, seed=55 + cps, heterogeneous=het)
    rng = np.random.default_rng(cps * 1000 + ncols)
    J = rng.uniform(-1.0, 1.0, (n, 6, ncols))
    base = rng.uniform(-50.0, 50.0, (ns, ncols)) if with_base else None
    planes_h = syn.aos_to_planes(st["twists"], st["poses"], st["null_poses"])
    prm_h = np.ascontiguousarray(st["params"].T) if het else None
    ref, wref = so.generalized_force(cps, ncols, planes_h, J, base, param_planes=prm_h,
                                     uniform=syn.REFERENCE_TEST_PARAMS, want_wrench=True,
                                     nthreads=NTHREADS)
    if aligned:
        Jd = _dev(torch, J)
    else:
        buf = torch.empty(J.size + 1, dtype=torch.float64, device="cuda")
        Jd = buf[1:].view(n, 6, ncols)
        Jd.copy_(torch.from_numpy(J))
    out, wr = gf.run(cps, ncols, _dev(torch, planes_h), Jd, _dev(torch, base),
                     param_planes=_dev(torch, prm_h), want_wrench=True)
    assert batch.handle.last_path == (1 if aligned else 2)
    Wm = np.abs(wref.T.reshape(ns, cps, 6))
    mag = (np.abs(base) if with_base else 0.0) + \
        np.einsum("scrq,scr->sq", np.abs(J.reshape(ns, cps, 6, ncols)), Wm)
    err = np.abs(out.cpu().numpy() - ref).max(axis=1) / np.maximum(mag.max(axis=1), 1e-300)
    assert err.max() <= TOL, (int(np.argmax(err)), err.max())
    assert_parity(wr.cpu().numpy().T, wref.T, "wrench")
    if with_base:   # in place: base aliases out
        b = _dev(torch, base)
        gf.run(cps, ncols, _dev(torch, planes_h), Jd, b, param_planes=_dev(torch, prm_h), out=b)
        assert torch.equal(b, out)


def test_generalized_force_argument_errors(torch, batch):
    from bipedal_locomotion_framework_b200 import _capi
    from bipedal_locomotion_framework_b200.system import GeneralizedForceBatch
    gf = GeneralizedForceBatch(batch)
    z = lambda *s: torch.zeros(s, dtype=torch.float64, device="cuda")
    with pytest.raises(_capi.BlfCcmError):
        gf.run(33, 4, z(30, 33), z(33, 6, 4))
    with pytest.raises(_capi.BlfCcmError):
        gf.run(1, 129, z(30, 2), z(2, 6, 129))


# --- per-contact parameter table from the on-disk configuration format -----------------------------

def test_parameter_table_from_ini_drives_the_batched_evaluation(torch, batch, oracle):
    from bipedal_locomotion_framework_b200.ini import load_ini_string
    n = 257
    st = syn.make_states(n, seed=66, heterogeneous=True)
    fmt = lambda col: ", ".join(repr(float(x)) for x in col)
    text = "[CONTACT_PARAMETERS]\n" + "".join(
        f"{key} ({fmt(st['params'][:, j])})\n"
        for j, key in enumerate(("length", "width", "spring_coeff", "damper_coeff")))
    table = batch.load_parameter_table(load_ini_string(text).getGroup("CONTACT_PARAMETERS"))
    assert table.shape == (4, n) and np.array_equal(table.cpu().numpy(), st["params"].T)  # repr round-trips
    planes = _dev(torch, syn.aos_to_planes(st["twists"], st["poses"], st["null_poses"]))
    out = batch.evaluate_soa(planes, table, 7)
    ref = oracle.eval_batch_states(st, mask=7, nthreads=1)
    assert_parity(out["wrench"].cpu().numpy().T, ref["wrench"], "wrench")
    assert_parity(out["autodyn"].cpu().numpy().T, ref["autodyn"], "autodyn")
    assert_parity(out["ctrl"].cpu().numpy(), ref["ctrl"], "ctrl")
    assert batch.load_parameter_table(load_ini_string("length (0.1)\n")) is None


# --- memory safety and edge shapes ------------------------------------------------------------------

GUARD, SENTINEL = 64, -7.25


def _guarded(torch, rows, cols, fill):
    buf = torch.full((rows, cols + 2 * GUARD), fill, dtype=torch.float64, device="cuda")
    return buf, buf[:, GUARD:GUARD + cols]


def _check_guards(buf, cols, what):
    assert bool((buf[:, :GUARD] == SENTINEL).all()) and bool((buf[:, GUARD + cols:] == SENTINEL).all()), \
        f"{what}: guard band overwritten"


@pytest.mark.parametrize("nr,feet,H,het,rho", [(1, 1, 1, False, 0.0), (33, 1, 5, True, 3.0),
                                               (17, 3, 9, False, 0.5), (300, 2, 12, True, 0.0)])
def test_rollout_no_out_of_bounds_canaries(torch, batch, nr, feet, H, het, rho):
    """compute-sanitizer is closed on this pool: inputs are followed by NaN guards (an over-read
    would poison a result), every output lives between sentinel guard bands."""
    from bipedal_locomotion_framework_b200 import _capi
    import ctypes as C
    chains, n = nr * feet, nr * feet * H
    st = syn.make_states(chains, seed=5, heterogeneous=True)
    nan = float("nan")
    _, tw = _guarded(torch, 6, n, nan)
    tw.copy_(_dev(torch, syn.make_states(n, seed=6)["twists"].T))
    _, pos = _guarded(torch, 3, chains, nan)
    pos.copy_(_dev(torch, st["poses"][:, :3].T))
    _, rot = _guarded(torch, 9, chains, nan)
    rot.copy_(_dev(torch, st["poses"][:, 3:].T))
    _, nul = _guarded(torch, 12, chains, nan)
    nul.copy_(_dev(torch, st["null_poses"].T))
    _, prm = _guarded(torch, 4, chains, nan)
    prm.copy_(_dev(torch, st["params"].T))
    wb, w = _guarded(torch, 6, n, SENTINEL)
    ab, a = _guarded(torch, 6, n, SENTINEL)
    cb, c = _guarded(torch, 1, n * 36, SENTINEL)
    fpb, fp = _guarded(torch, 3, chains, SENTINEL)
    frb, fr = _guarded(torch, 9, chains, SENTINEL)
    costb, cost = _guarded(torch, 1, nr, SENTINEL)
    best = torch.empty(2, dtype=torch.int64, device="cuda")
    ref, wts = np.zeros(6), np.ones(2)
    pp = batch._plane_ptrs
    for mask in (7, 0):
        rc = _capi.lib().blf_ccm_rollout_integrate_cost(
            batch.handle.ptr, nr, feet, H, 0.01, rho, pp(tw, 6), pp(pos, 3), pp(rot, 9), pp(nul, 12),
            pp(prm, 4) if het else None, mask, pp(w, 6), pp(a, 6), c.data_ptr(), pp(fp, 3), pp(fr, 9),
            ref.ctypes.data_as(C.c_void_p), wts.ctypes.data_as(C.c_void_p), 0, cost.data_ptr(),
            best.data_ptr(), None)
        assert rc == 0, _capi.lib().blf_ccm_last_error()
        torch.cuda.synchronize()
        for b_, cols, what in ((wb, n, "wrench"), (ab, n, "autodyn"), (cb, n * 36, "ctrl"),
                               (fpb, chains, "final pos"), (frb, chains, "final rot"), (costb, nr, "cost")):
            _check_guards(b_, cols, f"rollout {what} mask={mask}")
        for t_ in (w, a, c, fp, fr, cost):
            assert bool(torch.isfinite(t_).all())            # no NaN guard was read


@pytest.mark.parametrize("ns,cps,ncols", [(1, 1, 1), (3, 2, 29), (11, 3, 33), (5, 32, 128), (40, 7, 64),
                                          (9, 2, 6), (13, 3, 12), (2, 32, 16)])
def test_generalized_force_no_out_of_bounds_canaries(torch, batch, ns, cps, ncols):
    from bipedal_locomotion_framework_b200.system import GeneralizedForceBatch
    gf = GeneralizedForceBatch(batch)
    n = ns * cps
    st = syn.make_states(n, seed=8)
    nan = float("nan")
    _, planes = _guarded(torch, 30, n, nan)
    planes.copy_(_dev(torch, syn.aos_to_planes(st["twists"], st["poses"], st["null_poses"])))
    _, J = _guarded(torch, 1, n * 6 * ncols, nan)
    J.copy_(torch.rand((1, n * 6 * ncols), dtype=torch.float64, device="cuda"))
    _, base = _guarded(torch, 1, ns * ncols, nan)
    base.copy_(torch.rand((1, ns * ncols), dtype=torch.float64, device="cuda"))
    ob, out = _guarded(torch, 1, ns * ncols, SENTINEL)
    res = gf.run(cps, ncols, planes, J.view(n, 6, ncols), base.view(ns, ncols), out=out.view(ns, ncols))
    torch.cuda.synchronize()
    _check_guards(ob, ns * ncols, "generalized force out")
    assert bool(torch.isfinite(res).all())


def test_euler_step_no_out_of_bounds_canaries(torch, batch):
    from bipedal_locomotion_framework_b200.system import KinematicsBatch
    kb = KinematicsBatch(0, batch.handle)
    for n in (1, 31, 129, 1000):
        st = syn.make_states(n, seed=9)
        _, tw = _guarded(torch, 6, n, float("nan"))
        tw.copy_(_dev(torch, st["twists"].T))
        pb, p = _guarded(torch, 3, n, SENTINEL)
        p.copy_(_dev(torch, st["poses"][:, :3].T))
        rb, r = _guarded(torch, 9, n, SENTINEL)
        r.copy_(_dev(torch, st["poses"][:, 3:].T))
        kb.euler_step(1.5, 1e-3, tw, p, r)
        torch.cuda.synchronize()
        _check_guards(pb, n, "euler pos")
        _check_guards(rb, n, "euler rot")
        assert bool(torch.isfinite(p).all()) and bool(torch.isfinite(r).all())


def test_rollout_forced_split_variants_agree_with_oracle(torch, so):
    """Every warps-per-tile variant (1, 2, 4, 8) of the rollout kernel against the oracle, with
    trajectories (forced through BLF_CCM_TUNE_ROLLOUT_SPLIT, read when a handle is created)."""
    from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch
    from bipedal_locomotion_framework_b200.system import RolloutBatch
    nr, feet, H = 45, 2, 23
    chains = nr * feet
    st = syn.make_states(chains, seed=12)
    tw = np.ascontiguousarray(syn.make_states(H * chains, seed=13)["twists"].T)
    pos0 = np.ascontiguousarray(st["poses"][:, :3].T)
    rot0 = np.ascontiguousarray(st["poses"][:, 3:].T)
    null = np.ascontiguousarray(st["null_poses"].T)
    ref_w, wts = np.array([0.0, 0.0, 30.0, 0.1, -0.1, 0.0]), np.array([1.0, 25.0])
    ref = so.rollout(nr, feet, H, 0.01, 1.0, tw, pos0, rot0, null, uniform=syn.REFERENCE_TEST_PARAMS,
                     mask=7, wrench_ref=ref_w, weights=wts)
    try:
        for split in (1, 2, 4, 8):
            os.environ["BLF_CCM_TUNE_ROLLOUT_SPLIT"] = str(split)
            b = ContinuousContactModelBatch(0)
            b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
            out = RolloutBatch(b).run(nr, feet, H, 0.01, 1.0, _dev(torch, tw), _dev(torch, pos0),
                                      _dev(torch, rot0), _dev(torch, null), ref_w, wts, mask=7,
                                      want_final=True)
            _check_traj(out, ref, 7, f"split={split}")
            assert rel(out["final_rot"].cpu().numpy().T, ref["rot"].T).max() <= TOL
            assert np.all(rel(out["cost"].cpu().numpy()[:, None], ref["cost"][:, None]) <= TOL)
            assert b.decode_best(out["best"])[1] == int(np.argmin(out["cost"].cpu().numpy()))
    finally:
        os.environ.pop("BLF_CCM_TUNE_ROLLOUT_SPLIT", None)


def test_rollout_large_batch_sampled_against_oracle(torch, batch, so):
    """13.1 M evaluations (65 536 rollouts x 2 feet x 100 steps, the size where the kernel is
    FP64-bound): every 257th rollout re-run by the oracle from the same device bits -- wrench
    trajectory, final pose and cost."""
    from bipedal_locomotion_framework_b200.system import RolloutBatch
    rb = RolloutBatch(batch)
    nr, feet, H = 65536, 2, 100
    chains = nr * feet
    g = torch.Generator(device="cuda")
    g.manual_seed(2024)
    U = lambda *s: torch.rand(s, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    tw = U(6, H * chains)
    base = syn.make_states(4096, seed=404)
    rep = chains // 4096
    pos = _dev(torch, np.tile(base["poses"][:, :3].T, (1, rep))) + 1e-3 * U(3, chains)
    rot = _dev(torch, np.tile(base["poses"][:, 3:].T, (1, rep))) * (1.0 + 1e-4 * U(9, chains))
    null = _dev(torch, np.tile(base["null_poses"].T, (1, rep)))
    ref_w, wts = np.array([0.0, 0.0, 30.0, 0.0, 0.0, 0.0]), np.array([1.0, 10.0])
    out = rb.run(nr, feet, H, 0.01, 0.01, tw, pos, rot, null, ref_w, wts, mask=1, want_final=True)
    rolls = torch.arange(0, nr, 257, device="cuda")
    cidx = (rolls[:, None] * feet + torch.arange(feet, device="cuda")[None, :]).reshape(-1)   # chains
    m = cidx.numel()
    tidx = (torch.arange(H, device="cuda")[:, None] * chains + cidx[None, :]).reshape(-1)     # (t, chain)
    ref = so.rollout(rolls.numel(), feet, H, 0.01, 0.01, tw[:, tidx].cpu().numpy(),
                     pos[:, cidx].cpu().numpy(), rot[:, cidx].cpu().numpy(),
                     null[:, cidx].cpu().numpy(), uniform=syn.REFERENCE_TEST_PARAMS, mask=1,
                     wrench_ref=ref_w, weights=wts, nthreads=NTHREADS)
    assert_parity(out["wrench"][:, tidx].cpu().numpy().T, ref["wrench"].T, "wrench", what="13M rollout ")
    assert rel(out["final_pos"][:, cidx].cpu().numpy().T, ref["pos"].T).max() <= TOL
    assert rel(out["final_rot"][:, cidx].cpu().numpy().T, ref["rot"].T).max() <= TOL
    got_cost = out["cost"][rolls].cpu().numpy()
    assert np.all(rel(got_cost[:, None], ref["cost"][:, None]) <= TOL)
    cost = out["cost"].cpu().numpy()
    c, idx = batch.decode_best(out["best"])
    assert idx == int(np.argmin(cost)) and c == cost.min()
    assert m == rolls.numel() * feet


def test_generalized_force_config2_size_sampled_against_oracle(torch, batch, so):
    """409 600 robots x 2 feet x 6x29 Jacobians (the configs[2] batch seen from the simulator side):
    every 1021st robot re-evaluated by the oracle from the same device bits."""
    from bipedal_locomotion_framework_b200.system import GeneralizedForceBatch
    gf = GeneralizedForceBatch(batch)
    ns, cps, ncols = 409600, 2, 29
    n = ns * cps
    st = syn.make_states(n, seed=45)
    planes = _dev(torch, syn.aos_to_planes(st["twists"], st["poses"], st["null_poses"]))
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    J = torch.rand((n, 6, ncols), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    base = torch.rand((ns, ncols), dtype=torch.float64, device="cuda", generator=g) * 100 - 50
    out = gf.run(cps, ncols, planes, J, base)
    sys_idx = torch.arange(0, ns, 1021, device="cuda")
    cidx = (sys_idx[:, None] * cps + torch.arange(cps, device="cuda")[None, :]).reshape(-1)
    Js, bs = J[cidx].cpu().numpy(), base[sys_idx].cpu().numpy()
    ref, wref = so.generalized_force(cps, ncols, planes[:, cidx].cpu().numpy(), Js, bs,
                                     uniform=syn.REFERENCE_TEST_PARAMS, want_wrench=True)
    m = sys_idx.numel()
    mag = np.abs(bs) + np.einsum("scrq,scr->sq", np.abs(Js.reshape(m, cps, 6, ncols)),
                                 np.abs(wref.T.reshape(m, cps, 6)))
    err = np.abs(out[sys_idx].cpu().numpy() - ref).max(axis=1) / mag.max(axis=1)
    assert err.max() <= TOL


@pytest.mark.parametrize("nr,feet,H,het,rho", [(5000, 2, 40, False, 0.0), (7, 3, 11, True, 2.0),
                                               (40000, 1, 100, True, 0.01)])
def test_rollout_host_pipeline_vs_device_path(torch, batch, nr, feet, H, het, rho):
    """blf_ccm_rollout_integrate_cost_host (chunked, three slots, strided H2D of the time-major
    twist planes) must reproduce the device-resident entry point bit for bit (same kernel, split
    forced to 1 there), whatever the chunking."""
    from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch
    from bipedal_locomotion_framework_b200.system import RolloutBatch
    chains = nr * feet
    st = syn.make_states(chains, seed=31, heterogeneous=het)
    tw = np.ascontiguousarray(syn.make_states(H * chains, seed=32)["twists"].T)
    pos0 = np.ascontiguousarray(st["poses"][:, :3].T)
    rot0 = np.ascontiguousarray(st["poses"][:, 3:].T)
    null = np.ascontiguousarray(st["null_poses"].T)
    prm = np.ascontiguousarray(st["params"].T) if het else None
    ref_w, wts = np.array([0.0, 0.0, 30.0, 0.1, -0.1, 0.0]), np.array([1.0, 25.0])
    os.environ["BLF_CCM_TUNE_ROLLOUT_SPLIT"] = "1"
    try:
        b1 = ContinuousContactModelBatch(0)
    finally:
        os.environ.pop("BLF_CCM_TUNE_ROLLOUT_SPLIT", None)
    b1.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
    dev = RolloutBatch(b1).run(nr, feet, H, 0.01, rho, _dev(torch, tw), _dev(torch, pos0),
                               _dev(torch, rot0), _dev(torch, null), ref_w, wts,
                               param_planes=_dev(torch, prm), mask=0)
    bc, bi, cost = RolloutBatch(batch).run_host(nr, feet, H, 0.01, rho, tw, pos0, rot0, null, ref_w,
                                                wts, param_planes=prm)
    assert np.array_equal(cost, dev["cost"].cpu().numpy())
    assert (bc, bi) == b1.decode_best(dev["best"])
    # pinned torch tensors work the same; the costs are optional
    pin = lambda a: None if a is None else torch.from_numpy(a).pin_memory()
    bc2, bi2, none = RolloutBatch(batch).run_host(nr, feet, H, 0.01, rho, pin(tw), pin(pos0), pin(rot0),
                                                  pin(null), ref_w, wts, param_planes=pin(prm),
                                                  want_cost=False)
    assert (bc2, bi2) == (bc, bi) and none is None
    assert RolloutBatch(batch).run_host(0, feet, H, 0.01, rho, tw[:, :0], pos0[:, :0], rot0[:, :0],
                                        null[:, :0], ref_w, wts)[1] == -1


# --- non-finite inputs and ties ---------------------------------------------------------------------

def test_non_finite_inputs_stay_local_and_never_win_the_argmin(torch, batch, so):
    """A NaN / Inf in one chain's state poisons that chain only (the evaluations are independent, as
    separate reference objects are); a rollout whose cost is NaN never wins the arg-min, exact ties go
    to the lowest index, and when every cost is NaN the index is -1."""
    from bipedal_locomotion_framework_b200.system import RolloutBatch
    rb = RolloutBatch(batch)
    nr, feet, H = 96, 2, 12
    chains = nr * feet
    st = syn.make_states(chains, seed=71)
    tw = np.ascontiguousarray(syn.make_states(H * chains, seed=72)["twists"].T)
    pos0 = np.ascontiguousarray(st["poses"][:, :3].T)
    rot0 = np.ascontiguousarray(st["poses"][:, 3:].T)
    null = np.ascontiguousarray(st["null_poses"].T)
    ref_w, wts = np.zeros(6), np.array([1.0, 1.0])
    clean = rb.run(nr, feet, H, 0.01, 0.5, _dev(torch, tw), _dev(torch, pos0), _dev(torch, rot0),
                   _dev(torch, null), ref_w, wts, mask=1, want_final=True)
    cost0 = clean["cost"].cpu().numpy()
    best0 = int(np.argmin(cost0))
    # poison the best rollout (NaN twist at step 3 of its first foot) and another one with Inf
    other = (best0 + 7) % nr
    tw2 = tw.copy()
    tw2[4, 3 * chains + best0 * feet] = np.nan
    pos2 = pos0.copy()
    pos2[1, other * feet + 1] = np.inf
    out = rb.run(nr, feet, H, 0.01, 0.5, _dev(torch, tw2), _dev(torch, pos2), _dev(torch, rot0),
                 _dev(torch, null), ref_w, wts, mask=1, want_final=True)
    cost = out["cost"].cpu().numpy()
    bad = np.zeros(nr, dtype=bool)
    bad[[best0, other]] = True
    assert not np.isfinite(cost[bad]).any() and np.array_equal(cost[~bad], cost0[~bad])
    good_chains = np.repeat(~bad, feet)
    assert torch.equal(out["final_rot"][:, torch.from_numpy(good_chains).cuda()],
                       clean["final_rot"][:, torch.from_numpy(good_chains).cuda()])
    c, idx = batch.decode_best(out["best"])
    expect = int(np.nanargmin(np.where(np.isfinite(cost), cost, np.nan)))
    assert idx == expect and c == cost[expect] and idx not in (best0, other)
    # exact tie: duplicate the data of rollout 5 into rollout 50 -> the lower index wins
    tw3, pos3, rot3, null3 = tw.copy(), pos0.copy(), rot0.copy(), null.copy()
    a, b = 5, 50
    for arr in (pos3, rot3, null3):
        arr[:, b * feet:(b + 1) * feet] = arr[:, a * feet:(a + 1) * feet]
    for t in range(H):
        tw3[:, t * chains + b * feet:t * chains + (b + 1) * feet] = \
            tw3[:, t * chains + a * feet:t * chains + (a + 1) * feet]
    # make that pair the cheapest: the reference wrench is their own mean wrench
    probe = rb.run(nr, feet, H, 0.01, 0.5, _dev(torch, tw3), _dev(torch, pos3), _dev(torch, rot3),
                   _dev(torch, null3), ref_w, wts, mask=1)
    w = probe["wrench"].cpu().numpy().reshape(6, H, chains)[:, :, a * feet:(a + 1) * feet]
    tie = rb.run(nr, feet, H, 0.01, 0.5, _dev(torch, tw3), _dev(torch, pos3), _dev(torch, rot3),
                 _dev(torch, null3), w.mean(axis=(1, 2)), wts, mask=0)
    ct = tie["cost"].cpu().numpy()
    assert ct[a] == ct[b]
    if np.argmin(ct) in (a, b):
        assert batch.decode_best(tie["best"])[1] == a
    # every cost NaN -> nothing comparable
    allnan = rb.run(nr, feet, H, 0.01, 0.5, _dev(torch, np.full_like(tw, np.nan)), _dev(torch, pos0),
                    _dev(torch, rot0), _dev(torch, null), ref_w, wts, mask=0)
    assert batch.decode_best(allnan["best"])[1] == -1


@pytest.mark.parametrize("variant", [33, 34, 35, 23, 25, 27, 13, 17, 3, 7])
@pytest.mark.parametrize("het,rho,nr,feet,H", [(True, 3.0, 100, 2, 31), (False, 0.0, 33, 1, 2),
                                               (True, 0.0, 10, 3, 10), (False, 0.5, 700, 2, 64),
                                               (False, 0.01, 4096, 2, 100), (False, 2.0, 37, 4, 17),
                                               (True, 0.01, 8, 2, 8), (False, 0.0, 5, 2, 1)])
def test_rollout_warp_specialised_variant_vs_oracle(torch, so, variant, het, rho, nr, feet, H):
    """The warp-specialised rollout kernels forced for every template instance, incl. horizons
    shorter than a TMA box / the stage ring, ragged tiles, an odd number of chains and a foot count
    that does not divide 32 (where the TMA forms must hand over to the second).  variant 33/34/35:
    ccm_rollout_ws5_kernel (hand-over per eight-step box, straight-line producer, consumers read the
    twists from the TMA box; warp layout 0/1/2 of Ws5Cfg: 3 or 4 consumer warps and a TMA loader warp; ONE launch); 23/25/27:
    ccm_rollout_ws4_kernel (four lanes per chain, TMA twists, fused reduction: ONE launch) with 3/5/7
    consumer warps; 13/17: ccm_rollout_ws3_kernel (one lane per chain, otherwise the same); 3/7:
    ccm_rollout_ws2_kernel + ccm_cost_reduce_kernel (two launches)."""
    from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch
    from bipedal_locomotion_framework_b200.system import RolloutBatch
    chains = nr * feet
    st = syn.make_states(chains, seed=88, heterogeneous=het)
    tw = np.ascontiguousarray(syn.make_states(H * chains, seed=89)["twists"].T)
    pos0 = np.ascontiguousarray(st["poses"][:, :3].T)
    rot0 = np.ascontiguousarray(st["poses"][:, 3:].T)
    null = np.ascontiguousarray(st["null_poses"].T)
    prm = np.ascontiguousarray(st["params"].T) if het else None
    ref_w, wts = np.array([0.0, 0.0, 30.0, 0.1, -0.1, 0.0]), np.array([1.0, 25.0])
    ref = so.rollout(nr, feet, H, 0.01, rho, tw, pos0, rot0, null, param_planes=prm,
                     uniform=syn.REFERENCE_TEST_PARAMS, mask=0, wrench_ref=ref_w, weights=wts)
    os.environ["BLF_CCM_TUNE_ROLLOUT_WS"] = str(variant)
    try:
        b = ContinuousContactModelBatch(0)
    finally:
        os.environ.pop("BLF_CCM_TUNE_ROLLOUT_WS", None)
    b.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
    launches = b.handle.launch_count
    out = RolloutBatch(b).run(nr, feet, H, 0.01, rho, _dev(torch, tw), _dev(torch, pos0),
                              _dev(torch, rot0), _dev(torch, null), ref_w, wts,
                              param_planes=_dev(torch, prm), mask=0, want_final=True)
    third_form = variant > 10 and 32 % feet == 0 and chains % 2 == 0
    assert b.handle.launch_count - launches == (1 if third_form else 2)
    cost = out["cost"].cpu().numpy()
    assert np.all(rel(cost[:, None], ref["cost"][:, None]) <= TOL)
    assert rel(out["final_pos"].cpu().numpy().T, ref["pos"].T).max() <= TOL
    assert rel(out["final_rot"].cpu().numpy().T, ref["rot"].T).max() <= TOL
    c, idx = b.decode_best(out["best"])
    assert idx == int(np.argmin(cost)) and c == cost.min()
    out2 = RolloutBatch(b).run(nr, feet, H, 0.01, rho, _dev(torch, tw), _dev(torch, pos0),
                               _dev(torch, rot0), _dev(torch, null), ref_w, wts,
                               param_planes=_dev(torch, prm), mask=0)
    assert np.array_equal(out2["cost"].cpu().numpy(), cost)     # deterministic
    # every kernel form integrates through the same kin_euler_step: the final poses are bit-identical
    # to the plain one-lane-per-chain kernel (the costs differ in the last bits: other summation order)
    os.environ["BLF_CCM_TUNE_ROLLOUT_WS"] = "2"
    try:
        bp = ContinuousContactModelBatch(0)
    finally:
        os.environ.pop("BLF_CCM_TUNE_ROLLOUT_WS", None)
    bp.set_uniform_params(*syn.REFERENCE_TEST_PARAMS)
    plain = RolloutBatch(bp).run(nr, feet, H, 0.01, rho, _dev(torch, tw), _dev(torch, pos0),
                                 _dev(torch, rot0), _dev(torch, null), ref_w, wts,
                                 param_planes=_dev(torch, prm), mask=0, want_final=True)
    assert torch.equal(plain["final_pos"], out["final_pos"]) and torch.equal(plain["final_rot"], out["final_rot"])
    assert np.all(rel(plain["cost"].cpu().numpy()[:, None], cost[:, None]) <= 1e-13)
    # index_base and a NULL cost array through the fused reduction
    out3 = RolloutBatch(b).run(nr, feet, H, 0.01, rho, _dev(torch, tw), _dev(torch, pos0),
                               _dev(torch, rot0), _dev(torch, null), ref_w, wts,
                               param_planes=_dev(torch, prm), mask=0, index_base=1000, want_cost=False)
    assert b.decode_best(out3["best"]) == (c, idx + 1000)


@pytest.mark.gpu
@pytest.mark.parametrize("rho", [0.0, 0.7])
def test_kinematics_dynamics_device_arrays_vs_oracle_and_host_call(torch, batch, so, rho):
    """blf_sys_kinematics_dynamics (device arrays, asynchronous) against the C oracle system by system
    and bit for bit against blf_sys_kinematics_dynamics_host (same kernel behind both)."""
    from bipedal_locomotion_framework_b200 import _capi
    L, h = _capi.lib(), batch.handle.ptr
    n = 1_003
    rng = np.random.default_rng(12)
    tw = rng.normal(size=(n, 6))
    R = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] * (1.0 + 0.02 * rng.normal()) for _ in range(n)])
    d_tw, d_R = _dev(torch, tw), _dev(torch, R.reshape(n, 9))
    d_pd = torch.full((n + 2, 3), 7.5, dtype=torch.float64, device="cuda")
    d_rd = torch.full((n + 2, 9), 7.5, dtype=torch.float64, device="cuda")
    assert L.blf_sys_kinematics_dynamics(h, n, rho, d_tw.data_ptr(), d_R.data_ptr(), d_pd[1:].data_ptr(),
                                         d_rd[1:].data_ptr(), None) == 0
    pd, rd = d_pd.cpu().numpy(), d_rd.cpu().numpy()
    assert (pd[0] == 7.5).all() and (pd[-1] == 7.5).all() and (rd[0] == 7.5).all() and (rd[-1] == 7.5).all()
    hp, hr = np.empty((n, 3)), np.empty((n, 9))
    ptr = lambda a: a.ctypes.data
    assert L.blf_sys_kinematics_dynamics_host(h, n, rho, ptr(tw), ptr(np.ascontiguousarray(R.reshape(n, 9))),
                                              ptr(hp), ptr(hr)) == 0
    assert np.array_equal(pd[1:-1], hp) and np.array_equal(rd[1:-1], hr)
    for i in range(0, n, 17):
        wp, wr = so.kinematics_dynamics(rho, tw[i], R[i])
        assert np.array_equal(hp[i], wp)
        assert rel(hr[i].reshape(3, 3), np.asarray(wr).reshape(3, 3)).max() <= TOL
    assert L.blf_sys_kinematics_dynamics(h, 1, rho, None, d_R.data_ptr(), d_pd.data_ptr(), d_rd.data_ptr(), None) != 0
    assert L.blf_sys_kinematics_dynamics(h, 0, rho, None, None, None, None, None) == 0

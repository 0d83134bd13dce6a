"""ctypes view of the System-component oracle (oracle/sys_oracle.c, inside libccm_oracle.so).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg, never from the product package.  Kinematics / integrator /
rollout / J^T-wrench accumulation pinned bit for bit against the reference's own sources compiled
into oracle/_ref (see sys_oracle.h for what the stand-ins are).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import ccm_oracle

_configured = False


def lib():
    global _configured
    L = ccm_oracle.lib()
    if not _configured:
        vp, dbl, ci, sz = C.c_void_p, C.c_double, C.c_int, C.c_size_t
        L.syso_kinematics_dynamics.argtypes = [dbl, vp, vp, vp, vp]
        L.syso_forward_euler_step.argtypes = [dbl, dbl, vp, vp, vp, ci, vp, vp]
        L.syso_integrate.argtypes = [dbl, dbl, dbl, dbl, vp, vp, vp, ci, vp, vp]
        L.syso_integrate.restype = ci
        L.syso_integrate_schedule.argtypes = [dbl, dbl, dbl, vp, ci]
        L.syso_integrate_schedule.restype = ci
        L.syso_euler_step_batch_soa.argtypes = [sz, dbl, dbl, vp, vp, vp, ci]
        L.syso_rollout.argtypes = [sz, ci, ci, dbl, dbl, vp, vp, vp, vp, vp, vp, C.c_uint, vp, vp,
                                   vp, vp, vp, vp, vp, ci]
        L.syso_generalized_force.argtypes = [sz, ci, ci, vp, vp, vp, vp, vp, vp, vp, ci]
        L.syso_mass_matrix_solve.argtypes = [sz, ci, vp, vp, vp, vp, vp, ci]
        L.syso_floating_base_euler_step.argtypes = [sz, ci, dbl, dbl, vp, vp, vp, vp, vp, ci]
        L.syso_floating_base_acceleration.argtypes = [sz, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                                      vp, ci]
        _configured = True
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _planes(arr2d):
    """(k, n) C-contiguous array -> ctypes array of k row pointers (None rows allowed via list)."""
    if arr2d is None:
        return None
    if isinstance(arr2d, (list, tuple)):
        return (C.c_void_p * len(arr2d))(*[None if r is None else r.ctypes.data for r in arr2d])
    assert arr2d.flags.c_contiguous and arr2d.dtype == np.float64
    return (C.c_void_p * arr2d.shape[0])(*[arr2d[i].ctypes.data for i in range(arr2d.shape[0])])


def kinematics_dynamics(rho, twist, rot):
    """FloatingBaseSystemKinematics::dynamics (base part).  rot (3,3) -> (pos_dot (3,), rot_dot (3,3))."""
    tw = np.ascontiguousarray(twist, dtype=np.float64).reshape(6)
    r = np.ascontiguousarray(rot, dtype=np.float64).reshape(9)
    pd, rd = np.empty(3), np.empty(9)
    lib().syso_kinematics_dynamics(float(rho), _ptr(tw), _ptr(r), _ptr(pd), _ptr(rd))
    return pd, rd.reshape(3, 3)


def forward_euler_step(rho, dT, twist, pos, rot, joint_vel=None, joint_pos=None):
    tw = np.ascontiguousarray(twist, dtype=np.float64).reshape(6)
    p = np.array(pos, dtype=np.float64).reshape(3).copy()
    r = np.array(rot, dtype=np.float64).reshape(9).copy()
    nj = 0 if joint_vel is None else len(joint_vel)
    jv = None if joint_vel is None else np.ascontiguousarray(joint_vel, dtype=np.float64)
    jp = None if joint_pos is None else np.array(joint_pos, dtype=np.float64).copy()
    lib().syso_forward_euler_step(float(rho), float(dT), _ptr(tw), _ptr(p), _ptr(r), nj, _ptr(jv),
                                  _ptr(jp))
    return (p, r.reshape(3, 3)) if joint_vel is None else (p, r.reshape(3, 3), jp)


def integrate(rho, step_dT, t0, tf, twist, pos, rot, joint_vel=None, joint_pos=None):
    """FixedStepIntegrator::integrate.  Returns (steps or -1, pos, rot[, joint_pos])."""
    tw = np.ascontiguousarray(twist, dtype=np.float64).reshape(6)
    p = np.array(pos, dtype=np.float64).reshape(3).copy()
    r = np.array(rot, dtype=np.float64).reshape(9).copy()
    nj = 0 if joint_vel is None else len(joint_vel)
    jv = None if joint_vel is None else np.ascontiguousarray(joint_vel, dtype=np.float64)
    jp = None if joint_pos is None else np.array(joint_pos, dtype=np.float64).copy()
    steps = lib().syso_integrate(float(rho), float(step_dT), float(t0), float(tf), _ptr(tw),
                                 _ptr(p), _ptr(r), nj, _ptr(jv), _ptr(jp))
    return (steps, p, r.reshape(3, 3)) if joint_vel is None else (steps, p, r.reshape(3, 3), jp)


def integrate_schedule(step_dT, t0, tf, cap=1 << 16):
    dts = np.empty(cap)
    c = lib().syso_integrate_schedule(float(step_dT), float(t0), float(tf), _ptr(dts), cap)
    return None if c < 0 else dts[:min(c, cap)].copy()


def euler_step_batch_soa(rho, dT, twist_planes, pos_planes, rot_planes, nthreads=1):
    """(6,n), (3,n), (9,n) -> new (pos_planes, rot_planes)."""
    tw = np.ascontiguousarray(twist_planes, dtype=np.float64)
    p = np.array(pos_planes, dtype=np.float64, order="C").copy()
    r = np.array(rot_planes, dtype=np.float64, order="C").copy()
    lib().syso_euler_step_batch_soa(tw.shape[1], float(rho), float(dT), _planes(tw), _planes(p),
                                    _planes(r), int(nthreads))
    return p, r


def rollout(n_rollouts, feet, horizon, dT, rho, twist_planes, pos_planes, rot_planes, null_planes,
            param_planes=None, uniform=None, mask=0, wrench_ref=None, weights=None, nthreads=1):
    """Fused rollout oracle.  twist_planes (6, horizon*chains) time-major; pos (3,chains),
    rot (9,chains), null (12,chains).  Returns dict with final pos/rot, per-mask trajectories,
    chain_cost, cost."""
    chains = n_rollouts * feet
    n = horizon * chains
    tw = np.ascontiguousarray(twist_planes, dtype=np.float64)
    assert tw.shape == (6, n)
    p = np.array(pos_planes, dtype=np.float64, order="C").copy()
    r = np.array(rot_planes, dtype=np.float64, order="C").copy()
    nu = np.ascontiguousarray(null_planes, dtype=np.float64)
    pr = None if param_planes is None else np.ascontiguousarray(param_planes, dtype=np.float64)
    uni = np.asarray(uniform if uniform is not None else (0, 0, 0, 0), dtype=np.float64)
    out = {
        "wrench": np.empty((6, n)) if mask & 1 else None,
        "autodyn": np.empty((6, n)) if mask & 2 else None,
        "ctrl": np.empty((n, 36)) if mask & 4 else None,
    }
    ref = None if wrench_ref is None else np.ascontiguousarray(wrench_ref, dtype=np.float64)
    wts = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    chain_cost = np.zeros(chains)
    cost = np.zeros(n_rollouts)
    lib().syso_rollout(n_rollouts, feet, horizon, float(dT), float(rho), _planes(tw), _planes(p),
                       _planes(r), _planes(nu), _planes(pr), _ptr(uni), mask,
                       _planes(out["wrench"]), _planes(out["autodyn"]), _ptr(out["ctrl"]),
                       _ptr(ref), _ptr(wts), _ptr(chain_cost), _ptr(cost), int(nthreads))
    out.update(pos=p, rot=r, chain_cost=chain_cost, cost=cost)
    return out


def generalized_force(contacts_per_system, ncols, in_planes, jacobians, base=None,
                      param_planes=None, uniform=None, want_wrench=False, nthreads=1):
    """out[s] = base[s] + sum_c J_c^T wrench_c.  in_planes (30,n) (or list with None for dead
    planes), jacobians (n, 6, ncols), base (n_systems, ncols) or None."""
    planes = in_planes
    n = (planes.shape[1] if not isinstance(planes, (list, tuple))
         else next(p for p in planes if p is not None).shape[0])
    assert n % contacts_per_system == 0
    ns = n // contacts_per_system
    J = np.ascontiguousarray(jacobians, dtype=np.float64)
    assert J.size == n * 6 * ncols
    b = None if base is None else np.ascontiguousarray(base, dtype=np.float64)
    pr = None if param_planes is None else np.ascontiguousarray(param_planes, dtype=np.float64)
    uni = np.asarray(uniform if uniform is not None else (0, 0, 0, 0), dtype=np.float64)
    out = np.empty((ns, ncols))
    wr = np.empty((6, n)) if want_wrench else None
    lib().syso_generalized_force(ns, int(contacts_per_system), int(ncols), _planes(planes),
                                 _planes(pr), _ptr(uni), _ptr(J), _ptr(b), _ptr(out), _planes(wr),
                                 int(nthreads))
    return (out, wr) if want_wrench else out


def mass_matrix_solve(mass, known, joint_torques=None, reg=None, nthreads=1):
    """acc[s] = (mass[s] + reg).llt().solve(known[s] (+ joint torques on the tail)).
    mass (n, nc, nc), known (n, nc), joint_torques (n, nc-6) or None, reg (nc, nc) or None."""
    M = np.ascontiguousarray(mass, dtype=np.float64)
    k = np.ascontiguousarray(known, dtype=np.float64)
    ns, nc = k.shape
    assert M.shape == (ns, nc, nc)
    tau = None if joint_torques is None else np.ascontiguousarray(joint_torques, dtype=np.float64)
    rg = None if reg is None else np.ascontiguousarray(reg, dtype=np.float64)
    acc = np.empty((ns, nc))
    lib().syso_mass_matrix_solve(ns, int(nc), _ptr(M), _ptr(rg), _ptr(k), _ptr(tau), _ptr(acc),
                                 int(nthreads))
    return acc


def floating_base_acceleration(contacts_per_system, in_planes, jacobians, bias, mass,
                               joint_torques=None, reg=None, param_planes=None, uniform=None,
                               want_wrench=False, nthreads=1):
    """dynamics() from the bias forces on: (mass + reg).llt().solve(-bias + sum J^T wrench (+ tau))."""
    planes = in_planes
    n = (planes.shape[1] if not isinstance(planes, (list, tuple))
         else next(p for p in planes if p is not None).shape[0])
    assert n % contacts_per_system == 0
    ns = n // contacts_per_system
    b = np.ascontiguousarray(bias, dtype=np.float64)
    nc = b.shape[1]
    J = np.ascontiguousarray(jacobians, dtype=np.float64)
    assert J.size == n * 6 * nc
    M = np.ascontiguousarray(mass, dtype=np.float64)
    assert M.shape == (ns, nc, nc)
    tau = None if joint_torques is None else np.ascontiguousarray(joint_torques, dtype=np.float64)
    rg = None if reg is None else np.ascontiguousarray(reg, dtype=np.float64)
    pr = None if param_planes is None else np.ascontiguousarray(param_planes, dtype=np.float64)
    uni = np.asarray(uniform if uniform is not None else (0, 0, 0, 0), dtype=np.float64)
    acc = np.empty((ns, nc))
    wr = np.empty((6, n)) if want_wrench else None
    lib().syso_floating_base_acceleration(ns, int(contacts_per_system), int(nc), _planes(planes),
                                          _planes(pr), _ptr(uni), _ptr(J), _ptr(b), _ptr(tau), _ptr(M),
                                          _ptr(rg), _ptr(acc), _planes(wr), int(nthreads))
    return (acc, wr) if want_wrench else acc


def floating_base_euler_step(rho, dT, acc, nu, joint_pos, base_pos, base_rot, nthreads=1):
    """One ForwardEuler step of FloatingBaseDynamicalSystem on copies of the state arrays:
    nu (n, nc), joint_pos (n, nc-6) or None, base_pos (n, 3), base_rot (n, 3, 3) -> the new four."""
    a = np.ascontiguousarray(acc, dtype=np.float64)
    v = np.array(nu, dtype=np.float64).copy()
    ns, nc = v.shape
    jp = None if joint_pos is None else np.array(joint_pos, dtype=np.float64).copy()
    p = np.array(base_pos, dtype=np.float64).copy()
    r = np.array(base_rot, dtype=np.float64).reshape(ns, 9).copy()
    lib().syso_floating_base_euler_step(ns, int(nc), float(rho), float(dT), _ptr(a), _ptr(v), _ptr(jp), _ptr(p),
                                        _ptr(r), int(nthreads))
    return v, jp, p, r.reshape(ns, 3, 3)

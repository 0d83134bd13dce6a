"""ctypes view of oracle/libccm_oracle.so (the C restatement in ccm_oracle.c).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg, never from the product package.  Pinned bit for bit against the
reference's own sources compiled into oracle/_ref (oracle/ref_binding.py); not against a binary linked
with the real Eigen/iDynTree (see ccm_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libccm_oracle.so")

WRENCH, AUTODYN, CTRL, REGRESSOR = 1, 2, 4, 8


def build(force: bool = False) -> str:
    """Compile the C oracle with oracle/Makefile (gcc -O2 -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("ccm_oracle.c", "ccm_oracle.h", "rls_oracle.c",
                                            "rls_oracle.h", "sys_oracle.c", "sys_oracle.h",
                                            "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in src))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libccm_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


class _Model(C.Structure):
    _fields_ = [
        ("is_wrench_computed", C.c_int), ("is_autodyn_computed", C.c_int),
        ("is_ctrl_computed", C.c_int), ("is_regressor_computed", C.c_int),
        ("wrench", C.c_double * 6), ("autodyn", C.c_double * 6),
        ("ctrl", C.c_double * 36), ("regressor", C.c_double * 12),
        ("frame", C.c_double * 12), ("null_force", C.c_double * 12),
        ("twist", C.c_double * 6),
        ("spring", C.c_double), ("damper", C.c_double),
        ("length", C.c_double), ("width", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.ccmo_get_contact_wrench.restype = dp
        L.ccmo_get_autonomous_dynamics.restype = dp
        L.ccmo_get_control_matrix.restype = dp
        L.ccmo_get_regressor.restype = dp
        L.ccmo_now.restype = C.c_double
        L.ccmo_initialize.argtypes = [C.c_void_p] + [C.c_double] * 4
        L.ccmo_get_force_at_point.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.ccmo_get_torque_generated_at_point.argtypes = [C.c_void_p, C.c_double, C.c_double,
                                                         C.c_void_p]
        L.ccmo_eval_batch_aos.argtypes = [C.c_size_t] + [C.c_void_p] * 5 + [C.c_uint] + \
            [C.c_void_p] * 4 + [C.c_int]
        L.ccmo_eval_batch_soa.argtypes = [C.c_size_t] + [C.c_void_p] * 3 + [C.c_uint] + \
            [C.c_void_p] * 4 + [C.c_int]
        L.ccmo_rollout_cost.argtypes = [C.c_size_t, C.c_size_t] + [C.c_void_p] * 4
        L.rlso_advance_batch.argtypes = [C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ContinuousContactModel:
    """Per-instance oracle object with the reference's method names
    (ContactModel.h:110-144, ContinuousContactModel.h:105-142)."""

    def __init__(self):
        self._m = _Model()
        lib().ccmo_construct(C.byref(self._m))

    def initialize(self, params: dict) -> bool:
        # ContinuousContactModel.cpp:35-57: four required double keys
        try:
            vals = [params[k] for k in ("length", "width", "spring_coeff", "damper_coeff")]
        except KeyError:
            return False
        if not all(isinstance(v, float) for v in vals):
            return False  # strict std::any_cast<double>, StdImplementation.tpp:33-42
        lib().ccmo_initialize(C.byref(self._m), *vals)
        return True

    def setState(self, twist, transform):
        tw = np.ascontiguousarray(twist, dtype=np.float64).reshape(6)
        tf = np.ascontiguousarray(transform, dtype=np.float64).reshape(12)
        lib().ccmo_set_state(C.byref(self._m), _ptr(tw), _ptr(tf))

    def setNullForceTransform(self, transform):
        tf = np.ascontiguousarray(transform, dtype=np.float64).reshape(12)
        lib().ccmo_set_null_force_transform(C.byref(self._m), _ptr(tf))

    def getContactWrench(self):
        return np.array(lib().ccmo_get_contact_wrench(C.byref(self._m))[0:6])

    def getAutonomousDynamics(self):
        return np.array(lib().ccmo_get_autonomous_dynamics(C.byref(self._m))[0:6])

    def getControlMatrix(self):
        return np.array(lib().ccmo_get_control_matrix(C.byref(self._m))[0:36]).reshape(6, 6)

    def getRegressor(self):
        return np.array(lib().ccmo_get_regressor(C.byref(self._m))[0:12]).reshape(6, 2)

    def getForceAtPoint(self, x, y):
        out = np.empty(3)
        lib().ccmo_get_force_at_point(C.byref(self._m), float(x), float(y), _ptr(out))
        return out

    def getTorqueGeneratedAtPoint(self, x, y):
        out = np.empty(3)
        lib().ccmo_get_torque_generated_at_point(C.byref(self._m), float(x), float(y), _ptr(out))
        return out

    # ContinuousContactModel.cpp:256-274 -- writing does not clear the lazy flags
    @property
    def springCoeff(self):
        return self._m.spring

    @springCoeff.setter
    def springCoeff(self, v):
        self._m.spring = v

    @property
    def damperCoeff(self):
        return self._m.damper

    @damperCoeff.setter
    def damperCoeff(self, v):
        self._m.damper = v


def eval_batch_aos(twists, poses, null_poses, params=None, uniform=None,
                   mask=WRENCH | AUTODYN | CTRL, nthreads=1) -> dict:
    """Per-instance path over a batch; returns dict of AoS outputs for the bits in mask."""
    n = twists.shape[0]
    twists = np.ascontiguousarray(twists, dtype=np.float64)
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    null_poses = np.ascontiguousarray(null_poses, dtype=np.float64)
    if params is not None:
        params = np.ascontiguousarray(params, dtype=np.float64)
    uni = np.asarray(uniform if uniform is not None else (0, 0, 0, 0), dtype=np.float64)
    out = {
        "wrench": np.empty((n, 6)) if mask & WRENCH else None,
        "autodyn": np.empty((n, 6)) if mask & AUTODYN else None,
        "ctrl": np.empty((n, 36)) if mask & CTRL else None,
        "regressor": np.empty((n, 12)) if mask & REGRESSOR else None,
    }
    lib().ccmo_eval_batch_aos(n, _ptr(twists), _ptr(poses), _ptr(null_poses), _ptr(params),
                              _ptr(uni), mask, _ptr(out["wrench"]), _ptr(out["autodyn"]),
                              _ptr(out["ctrl"]), _ptr(out["regressor"]), int(nthreads))
    return out


def eval_batch_states(states: dict, mask=WRENCH | AUTODYN | CTRL, nthreads=1) -> dict:
    return eval_batch_aos(states["twists"], states["poses"], states["null_poses"],
                          states.get("params"), states.get("uniform"), mask, nthreads)


def rollout_cost(wrench, rollout_len, wrench_ref, weights):
    wrench = np.ascontiguousarray(wrench, dtype=np.float64)
    n = wrench.shape[0]
    assert n % rollout_len == 0
    ref = np.ascontiguousarray(wrench_ref, dtype=np.float64)
    wts = np.ascontiguousarray(weights, dtype=np.float64)
    cost = np.empty(n // rollout_len)
    lib().ccmo_rollout_cost(n // rollout_len, rollout_len, _ptr(wrench), _ptr(ref), _ptr(wts),
                            _ptr(cost))
    return cost


def now() -> float:
    return lib().ccmo_now()


# --------------------------------------------------------------------------------------------------
# Estimators::RecursiveLeastSquare oracle (rls_oracle.c)
# --------------------------------------------------------------------------------------------------

def rls_advance_batch(Y, z, measurement_cov, lam, theta, P):
    """One advance() of n independent estimators.  Y (n,m,p), z (n,m), theta (n,p), P (n,p,p);
    returns new (theta, P) copies."""
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    n, m, p = Y.shape
    z = np.ascontiguousarray(z, dtype=np.float64).reshape(n, m)
    r = np.ascontiguousarray(measurement_cov, dtype=np.float64).reshape(m)
    theta = np.array(theta, dtype=np.float64).reshape(n, p).copy()
    P = np.array(P, dtype=np.float64).reshape(n, p, p).copy()
    lib().rlso_advance_batch(n, p, m, _ptr(Y), _ptr(z), _ptr(r), float(lam), _ptr(theta), _ptr(P))
    return theta, P


class RecursiveLeastSquare:
    """Per-instance oracle with the reference's method names
    (src/Estimators/include/BipedalLocomotion/Estimators/RecursiveLeastSquare.h:28-111)."""

    def __init__(self):
        self._ready = False
        self._regressor = None

    def initialize(self, params: dict) -> bool:
        if self._ready:
            return False  # "already initialized"
        try:
            self._r = np.array(params["measurement_covariance"], dtype=np.float64)
            self._lam = float(params["lambda"])
            self._theta = np.array(params["state"], dtype=np.float64)
            self._P = np.diag(np.array(params["state_covariance"], dtype=np.float64))
        except KeyError:
            return False
        self._z = np.zeros(self._r.size)
        self._ready = True
        return True

    def setRegressorFunction(self, fn):
        self._regressor = fn

    def setMeasurements(self, z):
        self._z = np.array(z, dtype=np.float64)

    def advance(self) -> bool:
        if self._regressor is None or not self._ready:
            return False
        Y = np.array(self._regressor(), dtype=np.float64)
        th, P = rls_advance_batch(Y[None], self._z[None], self._r, self._lam, self._theta[None],
                                  self._P[None])
        self._theta, self._P = th[0], P[0]
        return True

    def parametersExpectedValue(self):
        return self._theta

    def parametersCovarianceMatrix(self):
        return self._P

"""Python mirror of Estimators::RecursiveLeastSquare for the test / bench harness, computing only
through the C ABI (blf_rls_advance_*).  Reference:
src/Estimators/include/BipedalLocomotion/Estimators/RecursiveLeastSquare.h:28-111,
src/Estimators/src/RecursiveLeastSquare.cpp:17-149."""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _capi
from .contact_models import ContinuousContactModelBatch, _Handle, _np_ptr


class RecursiveLeastSquare:
    """Per-instance facade: same method names and state machine as the reference
    (NotInitialized -> Initialized -> Running); advance() is one n = 1 GPU evaluation."""

    def __init__(self, device: int = 0):
        self._device = device
        self._state = "NotInitialized"
        self._regressor = None
        self._handle = None

    def initialize(self, handler) -> bool:
        if self._state != "NotInitialized":
            print("[RecursiveLeastSquare::initialize] The estimator has been already initialized.",
                  file=sys.stderr)
            return False
        if handler is None:
            print("[RecursiveLeastSquare::initialize] The parameter handler is expired. Please check "
                  "its scope.", file=sys.stderr)
            return False
        ok, r = handler.getParameter("measurement_covariance", list)
        if not ok:
            print("[RecursiveLeastSquare::initialize] Unable to find the covariance matrix of the "
                  "measuraments.", file=sys.stderr)
            return False
        ok, lam = handler.getParameter("lambda", float)
        if not ok:
            print("[RecursiveLeastSquare::initialize] Unable to find lambda.", file=sys.stderr)
            return False
        ok, state = handler.getParameter("state", list)
        if not ok:
            print("[RecursiveLeastSquare::initialize] Unable to get the initial guess.",
                  file=sys.stderr)
            return False
        ok, scov = handler.getParameter("state_covariance", list)
        if not ok:
            print("[RecursiveLeastSquare::initialize] Unable to get the initial state covariance.",
                  file=sys.stderr)
            return False
        try:
            self._handle = _Handle(self._device)
        except (_capi.BlfCcmError, ImportError) as e:
            print(f"[RecursiveLeastSquare::initialize] CUDA backend unavailable: {e}", file=sys.stderr)
            return False
        self._r = np.array(r, dtype=np.float64)
        self._lambda = lam
        self._theta = np.array(state, dtype=np.float64)
        self._P = np.diag(np.array(scov, dtype=np.float64))
        self._z = np.zeros(self._r.size)
        self._state = "Initialized"
        return True

    def setRegressorFunction(self, fn) -> None:
        self._regressor = fn

    def setMeasurements(self, z) -> None:
        z = np.asarray(z, dtype=np.float64)
        assert z.size == self._z.size
        self._z = z.copy()

    def advance(self) -> bool:
        if self._regressor is None:
            print("[RecursiveLeastSquare::advance] Please call the setRegressorFunction() before "
                  "calling advance", file=sys.stderr)
            return False
        if self._state not in ("Initialized", "Running"):
            print("[RecursiveLeastSquare::advance] Please initialize the estimator before calling "
                  "advance.", file=sys.stderr)
            return False
        self._state = "Running"
        Y = np.ascontiguousarray(self._regressor(), dtype=np.float64)
        m, p = Y.shape
        _capi.check(_capi.lib().blf_rls_advance_host(
            self._handle.ptr, 1, p, m, _np_ptr(Y), _np_ptr(self._z), _np_ptr(self._r),
            self._lambda, _np_ptr(self._theta), _np_ptr(self._P)))
        return True

    def parametersExpectedValue(self):
        return self._theta

    def parametersCovarianceMatrix(self):
        return self._P


class RecursiveLeastSquareBatch:
    """n independent estimators on the device (SoA planes as torch CUDA tensors)."""

    def __init__(self, batch: ContinuousContactModelBatch, measurement_covariance, lam: float):
        self._b = batch
        self._r = np.ascontiguousarray(measurement_covariance, dtype=np.float64)
        self._lambda = float(lam)

    def advance(self, regressor_planes, measurement_planes, state_planes, cov_planes):
        """regressor (m*p, n), measurements (m, n), state (p, n) in/out, cov (p*p, n) in/out."""
        m, p = measurement_planes.shape[0], state_planes.shape[0]
        n = state_planes.shape[1]
        pp = self._b._plane_ptrs
        _capi.check(_capi.lib().blf_rls_advance_batch(
            self._b.handle.ptr, n, p, m, pp(regressor_planes, m * p), pp(measurement_planes, m),
            _np_ptr(self._r), self._lambda, pp(state_planes, p), pp(cov_planes, p * p),
            self._b._stream()))

    def prepare_advance(self, regressor_planes, measurement_planes, state_planes, cov_planes):
        """Pre-bound form of advance(): returns a zero-argument callable (launch-bound loops)."""
        m, p = measurement_planes.shape[0], state_planes.shape[0]
        n = state_planes.shape[1]
        pp = self._b._plane_ptrs
        args = (self._b.handle.ptr, n, p, m, pp(regressor_planes, m * p), pp(measurement_planes, m),
                _np_ptr(self._r), self._lambda, pp(state_planes, p), pp(cov_planes, p * p),
                self._b._stream())
        fn = _capi.lib().blf_rls_advance_batch
        keep = (regressor_planes, measurement_planes, state_planes, cov_planes)

        def call(_keep=keep):
            rc = fn(*args)
            if rc:
                _capi.check(rc)
        return call

    def prepare_advance_contacts(self, planes, measured_wrench_planes, state_planes, cov_planes,
                                 geometry_planes=None):
        """Pre-bound form of advance_contacts()."""
        n = self._b._num_contacts(planes)
        pp = self._b._plane_ptrs
        args = (self._b.handle.ptr, n, pp(planes, 30), pp(geometry_planes, 2),
                pp(measured_wrench_planes, 6), _np_ptr(self._r), self._lambda, pp(state_planes, 2),
                pp(cov_planes, 4), self._b._stream())
        fn = _capi.lib().blf_ccm_rls_advance_contacts
        keep = (planes, measured_wrench_planes, state_planes, cov_planes, geometry_planes)

        def call(_keep=keep):
            rc = fn(*args)
            if rc:
                _capi.check(rc)
        return call

    def advance_contacts(self, planes, measured_wrench_planes, state_planes, cov_planes,
                         geometry_planes=None):
        """Fused: regressor from the contact state in registers + one RLS step on (spring, damper)."""
        n = self._b._num_contacts(planes)
        pp = self._b._plane_ptrs
        _capi.check(_capi.lib().blf_ccm_rls_advance_contacts(
            self._b.handle.ptr, n, pp(planes, 30), pp(geometry_planes, 2),
            pp(measured_wrench_planes, 6), _np_ptr(self._r), self._lambda, pp(state_planes, 2),
            pp(cov_planes, 4), self._b._stream()))

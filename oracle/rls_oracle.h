/*
 * rls_oracle.h -- CPU oracle for Estimators::RecursiveLeastSquare.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as ccm_oracle.h).  PARITY STATUS: pinned against the
 * reference's own RecursiveLeastSquare.cpp compiled unmodified into oracle/_ref against stand-in
 * Eigen/iDynTree headers (bit-for-bit agreement required, tests/test_reference_build.py) -- with one
 * caveat: the dynamic-size inverse() underneath is the STAND-IN's partial-pivot LU, a restatement of
 * the algorithm Eigen documents, so the inverse itself stays unpinned against the real Eigen.
 * The reference's test (src/Estimators/tests/RecursiveLeastSquareTest.cpp:91-142) needs YARP and
 * holds no golden vectors, only "10 000 steps recover (43.2, 12.2) within 0.1 %".  Also pinned by
 * exact rational evaluation of the same update (oracle/exact_golden.py) and by that property.
 *
 * Follows src/Estimators/src/RecursiveLeastSquare.cpp:96-133 (advance) and :17-88 (initialize:
 * diagonal measurement covariance, lambda, initial state, diagonal state covariance):
 *   K     = P Y^T (lambda R + Y P Y^T)^-1          general inverse (Eigen dynamic .inverse() =
 *                                                  partial-pivot LU) -- restated as such
 *   theta = theta + K (z - Y theta)
 *   P     = (P - K Y P) / lambda
 * Y is m x p row-major, P is p x p row-major, R = diag(r) m x m.
 */
#ifndef RLS_ORACLE_H
#define RLS_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { RLSO_MAX_P = 8, RLSO_MAX_M = 8 };

typedef struct {
    int p, m;
    double lambda;
    double r[RLSO_MAX_M];                    /* diagonal of the measurement covariance */
    double theta[RLSO_MAX_P];                /* m_state */
    double P[RLSO_MAX_P * RLSO_MAX_P];       /* m_stateCovarianceMatrix, row-major p x p */
    double z[RLSO_MAX_M];                    /* m_measurements */
    double K[RLSO_MAX_P * RLSO_MAX_M];       /* m_kalmanGain, row-major p x m */
} rlso_estimator;

/* initialize(): state covariance = diag(state_cov), measurements zeroed */
int rlso_initialize(rlso_estimator* e, int p, int m, const double* measurement_cov, double lambda,
                    const double* state, const double* state_cov);
void rlso_set_measurements(rlso_estimator* e, const double* z);
/* advance() with the regressor Y (m x p row-major) the reference obtains from its callback */
int rlso_advance(rlso_estimator* e, const double* Y);

/* n independent estimators, AoS: Y n*(m*p), z n*m, theta n*p (in/out), P n*(p*p) (in/out) */
void rlso_advance_batch(size_t n, int p, int m, const double* Y, const double* z,
                        const double* measurement_cov, double lambda, double* theta, double* P);

#ifdef __cplusplus
}
#endif
#endif

/* rls_oracle.c -- see rls_oracle.h.  TEST INFRASTRUCTURE ONLY; parity status in rls_oracle.h. */
#include "rls_oracle.h"

#include <math.h>
#include <string.h>

int rlso_initialize(rlso_estimator* e, int p, int m, const double* measurement_cov, double lambda,
                    const double* state, const double* state_cov)
{
    if (p < 1 || p > RLSO_MAX_P || m < 1 || m > RLSO_MAX_M) return -1;
    memset(e, 0, sizeof(*e));
    e->p = p;
    e->m = m;
    e->lambda = lambda;
    for (int i = 0; i < m; ++i) e->r[i] = measurement_cov[i];
    for (int i = 0; i < p; ++i) {
        e->theta[i] = state[i];
        e->P[i * p + i] = state_cov[i];
    }
    return 0;
}

void rlso_set_measurements(rlso_estimator* e, const double* z)
{
    for (int i = 0; i < e->m; ++i) e->z[i] = z[i];
}

/* inverse by LU with partial pivoting (what Eigen's dynamic-size inverse() does) */
static int invert(int n, const double* a, double* inv)
{
    double lu[RLSO_MAX_M * RLSO_MAX_M];
    int perm[RLSO_MAX_M];
    memcpy(lu, a, sizeof(double) * (size_t)(n * n));
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = fabs(lu[k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (fabs(lu[i * n + k]) > best) {
                best = fabs(lu[i * n + k]);
                piv = i;
            }
        if (best == 0.0) return -1;
        if (piv != k) {
            for (int j = 0; j < n; ++j) {
                double t = lu[k * n + j];
                lu[k * n + j] = lu[piv * n + j];
                lu[piv * n + j] = t;
            }
            int t = perm[k];
            perm[k] = perm[piv];
            perm[piv] = t;
        }
        for (int i = k + 1; i < n; ++i) {
            lu[i * n + k] = lu[i * n + k] / lu[k * n + k];
            for (int j = k + 1; j < n; ++j) lu[i * n + j] = lu[i * n + j] - lu[i * n + k] * lu[k * n + j];
        }
    }
    for (int c = 0; c < n; ++c) {
        double y[RLSO_MAX_M];
        for (int i = 0; i < n; ++i) {
            double acc = (perm[i] == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; ++j) acc = acc - lu[i * n + j] * y[j];
            y[i] = acc;
        }
        for (int i = n - 1; i >= 0; --i) {
            double acc = y[i];
            for (int j = i + 1; j < n; ++j) acc = acc - lu[i * n + j] * inv[j * n + c];
            inv[i * n + c] = acc / lu[i * n + i];
        }
    }
    return 0;
}

int rlso_advance(rlso_estimator* e, const double* Y)
{
    const int p = e->p, m = e->m;
    double YP[RLSO_MAX_M * RLSO_MAX_P];   /* Y P       m x p */
    double S[RLSO_MAX_M * RLSO_MAX_M] = {0};  /* lambda R + (Y P) Y^T */
    double Sinv[RLSO_MAX_M * RLSO_MAX_M];
    double PYt[RLSO_MAX_P * RLSO_MAX_M];  /* P Y^T     p x m */

    for (int i = 0; i < m; ++i)
        for (int c = 0; c < p; ++c) {
            double acc = 0.0;
            for (int k = 0; k < p; ++k) acc = acc + Y[i * p + k] * e->P[k * p + c];
            YP[i * p + c] = acc;
        }
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            double acc = 0.0;
            for (int k = 0; k < p; ++k) acc = acc + YP[i * p + k] * Y[j * p + k];
            S[i * m + j] = (i == j ? e->lambda * e->r[i] : 0.0) + acc;
        }
    if (invert(m, S, Sinv) != 0) return -1;
    for (int r = 0; r < p; ++r)
        for (int j = 0; j < m; ++j) {
            double acc = 0.0;
            for (int k = 0; k < p; ++k) acc = acc + e->P[r * p + k] * Y[j * p + k];
            PYt[r * m + j] = acc;
        }
    for (int r = 0; r < p; ++r)
        for (int j = 0; j < m; ++j) {
            double acc = 0.0;
            for (int k = 0; k < m; ++k) acc = acc + PYt[r * m + k] * Sinv[k * m + j];
            e->K[r * m + j] = acc;
        }
    /* theta = theta + K (z - Y theta) */
    double innov[RLSO_MAX_M];
    for (int i = 0; i < m; ++i) {
        double acc = 0.0;
        for (int k = 0; k < p; ++k) acc = acc + Y[i * p + k] * e->theta[k];
        innov[i] = e->z[i] - acc;
    }
    for (int r = 0; r < p; ++r) {
        double acc = 0.0;
        for (int k = 0; k < m; ++k) acc = acc + e->K[r * m + k] * innov[k];
        e->theta[r] = e->theta[r] + acc;
    }
    /* P = (P - (K Y) P) / lambda */
    double KY[RLSO_MAX_P * RLSO_MAX_P], KYP[RLSO_MAX_P * RLSO_MAX_P];
    for (int r = 0; r < p; ++r)
        for (int c = 0; c < p; ++c) {
            double acc = 0.0;
            for (int k = 0; k < m; ++k) acc = acc + e->K[r * m + k] * Y[k * p + c];
            KY[r * p + c] = acc;
        }
    for (int r = 0; r < p; ++r)
        for (int c = 0; c < p; ++c) {
            double acc = 0.0;
            for (int k = 0; k < p; ++k) acc = acc + KY[r * p + k] * e->P[k * p + c];
            KYP[r * p + c] = acc;
        }
    for (int i = 0; i < p * p; ++i) e->P[i] = (e->P[i] - KYP[i]) / e->lambda;
    return 0;
}

void rlso_advance_batch(size_t n, int p, int m, const double* Y, const double* z,
                        const double* measurement_cov, double lambda, double* theta, double* P)
{
    rlso_estimator e;
    for (size_t i = 0; i < n; ++i) {
        memset(&e, 0, sizeof(e));
        e.p = p;
        e.m = m;
        e.lambda = lambda;
        for (int k = 0; k < m; ++k) e.r[k] = measurement_cov[k];
        memcpy(e.theta, theta + i * (size_t)p, sizeof(double) * (size_t)p);
        memcpy(e.P, P + i * (size_t)(p * p), sizeof(double) * (size_t)(p * p));
        rlso_set_measurements(&e, z + i * (size_t)m);
        rlso_advance(&e, Y + i * (size_t)(m * p));
        memcpy(theta + i * (size_t)p, e.theta, sizeof(double) * (size_t)p);
        memcpy(P + i * (size_t)(p * p), e.P, sizeof(double) * (size_t)(p * p));
    }
}

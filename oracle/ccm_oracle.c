/*
 * ccm_oracle.c -- CPU oracle for ContactModels::ContinuousContactModel.  TEST INFRASTRUCTURE ONLY,
 * parity pinned against the reference's own sources compiled into oracle/_ref, not against a binary
 * with the real Eigen (see ccm_oracle.h for both statements).
 *
 * The arithmetic keeps the reference's expression structure: explicit skew matrices, 3x3
 * matrix-matrix products associated left to right the way the C++ expressions parse, one rounding
 * per operation (build with -ffp-contract=off).  It is deliberately NOT the simplified
 * cross-product form the CUDA kernels use, so the two are independent derivations.
 */
#include "ccm_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- tiny fixed-size algebra, row-major 3x3 ------------------------------------------------- */

/* iDynTree::skew */
static void skew(const double v[3], double s[9])
{
    s[0] = 0.0;   s[1] = -v[2]; s[2] = v[1];
    s[3] = v[2];  s[4] = 0.0;   s[5] = -v[0];
    s[6] = -v[1]; s[7] = v[0];  s[8] = 0.0;
}

/* c = a * b, coefficient-wise sum over k = 0,1,2 in order */
static void mm(const double a[9], const double b[9], double c[9])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double acc = a[3 * i] * b[j];
            acc = acc + a[3 * i + 1] * b[3 + j];
            acc = acc + a[3 * i + 2] * b[6 + j];
            c[3 * i + j] = acc;
        }
}

static void mv(const double a[9], const double x[3], double y[3])
{
    for (int i = 0; i < 3; ++i) {
        double acc = a[3 * i] * x[0];
        acc = acc + a[3 * i + 1] * x[1];
        acc = acc + a[3 * i + 2] * x[2];
        y[i] = acc;
    }
}

static void sm(double s, const double a[9], double c[9])
{
    for (int i = 0; i < 9; ++i) c[i] = s * a[i];
}

static void madd(const double a[9], const double b[9], double c[9])
{
    for (int i = 0; i < 9; ++i) c[i] = a[i] + b[i];
}

static void col(const double r[9], int j, double c[3])
{
    c[0] = r[j]; c[1] = r[3 + j]; c[2] = r[6 + j];
}

/* ---- object protocol ------------------------------------------------------------------------ */

static void identity_transform(ccmo_transform* t)
{
    memset(t, 0, sizeof(*t));
    t->rot[0] = t->rot[4] = t->rot[8] = 1.0;
}

static void invalidate(ccmo_model* m)
{
    m->is_wrench_computed = 0;
    m->is_ctrl_computed = 0;
    m->is_autodyn_computed = 0;
    m->is_regressor_computed = 0;
}

void ccmo_construct(ccmo_model* m)
{
    /* ContinuousContactModel.cpp:16-22: control matrix, autonomous dynamics and regressor are
     * zeroed; the wrench is not (left as-is here too, but memset keeps valgrind quiet). */
    memset(m, 0, sizeof(*m));
    identity_transform(&m->frame);
    identity_transform(&m->null_force);
}

void ccmo_initialize(ccmo_model* m, double length, double width, double spring, double damper)
{
    invalidate(m);
    m->length = length;
    m->width = width;
    m->spring = spring;
    m->damper = damper;
}

void ccmo_set_state(ccmo_model* m, const ccmo_twist* twist, const ccmo_transform* transform)
{
    invalidate(m);
    m->twist = *twist;
    m->frame = *transform;
}

void ccmo_set_null_force_transform(ccmo_model* m, const ccmo_transform* transform)
{
    invalidate(m);
    m->null_force = *transform;
}

/* ---- ContinuousContactModel.cpp:79-108 ------------------------------------------------------ */

/* L*L*(b*S*S*w + k*S*n), parsed as ((L*L) * ((((b*S)*S)*w) + ((k*S)*n))) */
static void scaled_bracket_term(double len2, double b, double k, const double S[9],
                                const double w[3], const double n[3], double out[3])
{
    double bS[9], bSS[9], kS[9], t1[3], t2[3];
    sm(b, S, bS);
    mm(bS, S, bSS);
    mv(bSS, w, t1);
    sm(k, S, kS);
    mv(kS, n, t2);
    for (int i = 0; i < 3; ++i) out[i] = len2 * (t1[i] + t2[i]);
}

static void compute_contact_wrench(ccmo_model* m)
{
    const double area = m->length * m->width;
    const double* p = m->frame.pos;
    const double* R = m->frame.rot;
    const double* v = m->twist.lin;
    const double* w = m->twist.ang;
    const double* p0 = m->null_force.pos;
    const double* R0 = m->null_force.rot;
    const double k = m->spring, b = m->damper;

    /* :96-97  force = |R22| * area * (k*(p0 - p) - b*v) */
    const double s_f = fabs(R[8]) * area;
    for (int i = 0; i < 3; ++i)
        m->wrench[i] = s_f * (k * (p0[i] - p[i]) - b * v[i]);

    /* :100-107 */
    double e1[3], e2[3], n1[3], n2[3], S1[9], S2[9], a[3], c[3];
    col(R, 0, e1);
    col(R, 1, e2);
    col(R0, 0, n1);
    col(R0, 1, n2);
    skew(e1, S1);
    skew(e2, S2);
    scaled_bracket_term(m->length * m->length, b, k, S1, w, n1, a);
    scaled_bracket_term(m->width * m->width, b, k, S2, w, n2, c);
    const double s_t = fabs(R[8]) * area / 12;
    for (int i = 0; i < 3; ++i)
        m->wrench[3 + i] = s_t * (a[i] + c[i]);
}

/* ---- ContinuousContactModel.cpp:110-146 ----------------------------------------------------- */

/* L*L*(k*Sd*n + b*(Sd*S + S*Sd)*w) */
static void scaled_rate_term(double len2, double b, double k, const double S[9],
                             const double Sd[9], const double w[3], const double n[3],
                             double out[3])
{
    double kSd[9], t1[3], SdS[9], SSd[9], sum[9], bsum[9], t2[3];
    sm(k, Sd, kSd);
    mv(kSd, n, t1);
    mm(Sd, S, SdS);
    mm(S, Sd, SSd);
    madd(SdS, SSd, sum);
    sm(b, sum, bsum);
    mv(bsum, w, t2);
    for (int i = 0; i < 3; ++i) out[i] = len2 * (t1[i] + t2[i]);
}

static void compute_autonomous_dynamics(ccmo_model* m)
{
    const double area = m->length * m->width;
    const double* p = m->frame.pos;
    const double* R = m->frame.rot;
    const double* v = m->twist.lin;
    const double* w = m->twist.ang;
    const double* p0 = m->null_force.pos;
    const double* R0 = m->null_force.rot;
    const double k = m->spring, b = m->damper;

    /* :125  Rdot = skew(w) * R */
    double Sw[9], Rd[9];
    skew(w, Sw);
    mm(Sw, R, Rd);

    /* :127-129 (signed R22, no abs) */
    for (int i = 0; i < 3; ++i)
        m->autodyn[i] = area * (Rd[8] * (k * (p0[i] - p[i]) - b * v[i]) - R[8] * k * v[i]);

    /* :131-144 */
    double e1[3], e2[3], d1[3], d2[3], n1[3], n2[3], S1[9], S2[9], Sd1[9], Sd2[9];
    col(R, 0, e1);
    col(R, 1, e2);
    col(Rd, 0, d1);
    col(Rd, 1, d2);
    col(R0, 0, n1);
    col(R0, 1, n2);
    skew(e1, S1);
    skew(e2, S2);
    skew(d1, Sd1);
    skew(d2, Sd2);

    const double L2 = m->length * m->length, W2 = m->width * m->width;
    double a[3], c[3], ra[3], rc[3];
    scaled_bracket_term(L2, b, k, S1, w, n1, a);
    scaled_bracket_term(W2, b, k, S2, w, n2, c);
    scaled_rate_term(L2, b, k, S1, Sd1, w, n1, ra);
    scaled_rate_term(W2, b, k, S2, Sd2, w, n2, rc);
    const double s = area / 12;
    for (int i = 0; i < 3; ++i)
        m->autodyn[3 + i] = s * (Rd[8] * (a[i] + c[i]) + R[8] * (ra[i] + rc[i]));
}

/* ---- ContinuousContactModel.cpp:148-171 ----------------------------------------------------- */

/* (L*L*S1*S1 + W*W*S2*S2), each term parsed ((len2*S)*S) */
static void second_moment_matrix(double L2, double W2, const double S1[9], const double S2[9],
                                 double out[9])
{
    double t[9], a[9], c[9];
    sm(L2, S1, t);
    mm(t, S1, a);
    sm(W2, S2, t);
    mm(t, S2, c);
    madd(a, c, out);
}

static void compute_control_matrix(ccmo_model* m)
{
    const double area = m->length * m->width;
    const double* R = m->frame.rot;
    const double b = m->damper;

    /* :165 -- only the three diagonal entries of the top-left block are written */
    const double d = -area * b * R[8];
    m->ctrl[0] = d;
    m->ctrl[7] = d;
    m->ctrl[14] = d;

    /* :167-170 */
    double e1[3], e2[3], S1[9], S2[9], M[9];
    col(R, 0, e1);
    col(R, 1, e2);
    skew(e1, S1);
    skew(e2, S2);
    second_moment_matrix(m->length * m->length, m->width * m->width, S1, S2, M);
    const double s = area / 12 * R[8] * b;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            m->ctrl[6 * (3 + i) + 3 + j] = s * M[3 * i + j];
    /* every other entry keeps the +0.0 written by the constructor (:18) */
}

/* ---- ContinuousContactModel.cpp:223-254 ----------------------------------------------------- */

static void compute_regressor(ccmo_model* m)
{
    const double area = m->length * m->width;
    const double* p = m->frame.pos;
    const double* R = m->frame.rot;
    const double* v = m->twist.lin;
    const double* w = m->twist.ang;
    const double* p0 = m->null_force.pos;
    const double* R0 = m->null_force.rot;
    const double L2 = m->length * m->length, W2 = m->width * m->width;

    double e1[3], e2[3], n1[3], n2[3], S1[9], S2[9];
    col(R, 0, e1);
    col(R, 1, e2);
    col(R0, 0, n1);
    col(R0, 1, n2);
    skew(e1, S1);
    skew(e2, S2);

    const double s_f = fabs(R[8]) * area;
    const double s_v = -fabs(R[8]) * area;
    for (int i = 0; i < 3; ++i) {
        m->regressor[2 * i] = s_f * (p0[i] - p[i]);   /* top-left  :239-240 */
        m->regressor[2 * i + 1] = s_v * v[i];         /* top-right :242     */
    }

    /* bottom-left :244-247 */
    double t[9], a[3], c[3];
    sm(L2, S1, t);
    mv(t, n1, a);
    sm(W2, S2, t);
    mv(t, n2, c);
    const double s_t = area / 12.0 * fabs(R[8]);
    for (int i = 0; i < 3; ++i)
        m->regressor[2 * (3 + i)] = s_t * (a[i] + c[i]);

    /* bottom-right :249-253 : (s_t * M) * w */
    double M[9], sM[9], y[3];
    second_moment_matrix(L2, W2, S1, S2, M);
    sm(s_t, M, sM);
    mv(sM, w, y);
    for (int i = 0; i < 3; ++i)
        m->regressor[2 * (3 + i) + 1] = y[i];
}

/* ---- lazy getters, ContactModel.cpp:50-92 --------------------------------------------------- */

const double* ccmo_get_contact_wrench(ccmo_model* m)
{
    if (!m->is_wrench_computed) {
        compute_contact_wrench(m);
        m->is_wrench_computed = 1;
    }
    return m->wrench;
}

const double* ccmo_get_autonomous_dynamics(ccmo_model* m)
{
    if (!m->is_autodyn_computed) {
        compute_autonomous_dynamics(m);
        m->is_autodyn_computed = 1;
    }
    return m->autodyn;
}

const double* ccmo_get_control_matrix(ccmo_model* m)
{
    if (!m->is_ctrl_computed) {
        compute_control_matrix(m);
        m->is_ctrl_computed = 1;
    }
    return m->ctrl;
}

const double* ccmo_get_regressor(ccmo_model* m)
{
    if (!m->is_regressor_computed) {
        compute_regressor(m);
        m->is_regressor_computed = 1;
    }
    return m->regressor;
}

/* ---- ContinuousContactModel.cpp:173-221 ----------------------------------------------------- */

void ccmo_get_force_at_point(ccmo_model* m, double x, double y, double out[3])
{
    if (fabs(x) > m->length / 2 || fabs(y) > m->width / 2) {
        out[0] = out[1] = out[2] = 0.0;
        return;
    }
    const double* p = m->frame.pos;
    const double* R = m->frame.rot;
    const double* v = m->twist.lin;
    const double* w = m->twist.ang;
    const double* p0 = m->null_force.pos;
    const double* R0 = m->null_force.rot;
    const double q[3] = {x, y, 0.0};

    /* k*((p0 - p) + (R0 - R)*q) - b*(v + (skew(w)*R)*q)   :196-199 */
    double dR[9], dRq[3], Sw[9], SwR[9], SwRq[3];
    for (int i = 0; i < 9; ++i) dR[i] = R0[i] - R[i];
    mv(dR, q, dRq);
    skew(w, Sw);
    mm(Sw, R, SwR);
    mv(SwR, q, SwRq);
    for (int i = 0; i < 3; ++i)
        out[i] = m->spring * ((p0[i] - p[i]) + dRq[i]) - m->damper * (v[i] + SwRq[i]);
}

void ccmo_get_torque_generated_at_point(ccmo_model* m, double x, double y, double out[3])
{
    if (fabs(x) > m->length / 2 || fabs(y) > m->width / 2) {
        out[0] = out[1] = out[2] = 0.0;
        return;
    }
    const double q[3] = {x, y, 0.0};
    double r[3], f[3];
    mv(m->frame.rot, q, r);
    ccmo_get_force_at_point(m, x, y, f);
    /* (R*q).cross(force)   :219 */
    out[0] = r[1] * f[2] - r[2] * f[1];
    out[1] = r[2] * f[0] - r[0] * f[2];
    out[2] = r[0] * f[1] - r[1] * f[0];
}

/* ---- batch drivers -------------------------------------------------------------------------- */

typedef struct {
    int soa;
    size_t begin, end, n;
    /* AoS */
    const double *twists, *poses, *null_poses, *params;
    double *wrench, *autodyn, *ctrl, *regressor;
    /* SoA */
    const double* const* in_planes;
    const double* const* param_planes;
    double* const* wrench_planes;
    double* const* autodyn_planes;
    double* const* regressor_planes;
    const double* uniform;
    unsigned mask;
} job_t;

static void* run_job(void* arg)
{
    job_t* j = (job_t*)arg;
    ccmo_model m;
    ccmo_construct(&m);
    if (j->uniform)
        ccmo_initialize(&m, j->uniform[0], j->uniform[1], j->uniform[2], j->uniform[3]);
    for (size_t i = j->begin; i < j->end; ++i) {
        ccmo_twist tw;
        ccmo_transform tf, nf;
        if (j->soa) {
            const double* const* P = j->in_planes;
            for (int c = 0; c < 3; ++c) tw.lin[c] = P[c][i];
            for (int c = 0; c < 3; ++c) tw.ang[c] = P[3 + c][i];
            for (int c = 0; c < 3; ++c) tf.pos[c] = P[6 + c][i];
            for (int c = 0; c < 9; ++c) tf.rot[c] = P[9 + c][i];
            for (int c = 0; c < 3; ++c) nf.pos[c] = P[18 + c][i];
            for (int c = 0; c < 9; ++c) nf.rot[c] = P[21 + c] ? P[21 + c][i] : 0.0;
            if (j->param_planes)
                ccmo_initialize(&m, j->param_planes[0][i], j->param_planes[1][i],
                                j->param_planes[2][i], j->param_planes[3][i]);
        } else {
            memcpy(&tw, j->twists + 6 * i, sizeof(tw));
            memcpy(&tf, j->poses + 12 * i, sizeof(tf));
            memcpy(&nf, j->null_poses + 12 * i, sizeof(nf));
            if (j->params)
                ccmo_initialize(&m, j->params[4 * i], j->params[4 * i + 1], j->params[4 * i + 2],
                                j->params[4 * i + 3]);
        }
        ccmo_set_state(&m, &tw, &tf);
        ccmo_set_null_force_transform(&m, &nf);
        if (j->mask & CCMO_WRENCH) {
            const double* r = ccmo_get_contact_wrench(&m);
            if (j->soa) for (int c = 0; c < 6; ++c) j->wrench_planes[c][i] = r[c];
            else memcpy(j->wrench + 6 * i, r, 6 * sizeof(double));
        }
        if (j->mask & CCMO_AUTODYN) {
            const double* r = ccmo_get_autonomous_dynamics(&m);
            if (j->soa) for (int c = 0; c < 6; ++c) j->autodyn_planes[c][i] = r[c];
            else memcpy(j->autodyn + 6 * i, r, 6 * sizeof(double));
        }
        if (j->mask & CCMO_CTRL)
            memcpy(j->ctrl + 36 * i, ccmo_get_control_matrix(&m), 36 * sizeof(double));
        if (j->mask & CCMO_REGRESSOR) {
            const double* r = ccmo_get_regressor(&m);
            if (j->soa) for (int c = 0; c < 12; ++c) j->regressor_planes[c][i] = r[c];
            else memcpy(j->regressor + 12 * i, r, 12 * sizeof(double));
        }
    }
    return NULL;
}

static void dispatch(job_t* proto, size_t n, int nthreads)
{
    if (nthreads <= 1 || n < (size_t)nthreads) {
        proto->begin = 0;
        proto->end = n;
        run_job(proto);
        return;
    }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = *proto;
        jobs[t].begin = n * (size_t)t / (size_t)nthreads;
        jobs[t].end = n * (size_t)(t + 1) / (size_t)nthreads;
        pthread_create(&th[t], NULL, run_job, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(jobs);
    free(th);
}

void ccmo_eval_batch_aos(size_t n, const double* twists, const double* poses,
                         const double* null_poses, const double* params, const double uniform[4],
                         unsigned mask, double* wrench, double* autodyn, double* ctrl,
                         double* regressor, int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.soa = 0;
    j.n = n;
    j.twists = twists;
    j.poses = poses;
    j.null_poses = null_poses;
    j.params = params;
    j.uniform = params ? NULL : uniform;
    j.mask = mask;
    j.wrench = wrench;
    j.autodyn = autodyn;
    j.ctrl = ctrl;
    j.regressor = regressor;
    dispatch(&j, n, nthreads);
}

void ccmo_eval_batch_soa(size_t n, const double* const* in_planes,
                         const double* const* param_planes, const double uniform[4],
                         unsigned mask, double* const* wrench_planes,
                         double* const* autodyn_planes, double* ctrl,
                         double* const* regressor_planes, int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.soa = 1;
    j.n = n;
    j.in_planes = in_planes;
    j.param_planes = param_planes;
    j.uniform = param_planes ? NULL : uniform;
    j.mask = mask;
    j.wrench_planes = wrench_planes;
    j.autodyn_planes = autodyn_planes;
    j.ctrl = ctrl;
    j.regressor_planes = regressor_planes;
    dispatch(&j, n, nthreads);
}

void ccmo_rollout_cost(size_t n_rollouts, size_t rollout_len, const double* wrench,
                       const double wrench_ref[6], const double weights[2], double* cost)
{
    for (size_t r = 0; r < n_rollouts; ++r) {
        double acc = 0.0;
        for (size_t e = 0; e < rollout_len; ++e) {
            const double* w = wrench + 6 * (r * rollout_len + e);
            double qf = 0.0, qt = 0.0;
            for (int c = 0; c < 3; ++c) {
                const double df = w[c] - wrench_ref[c];
                const double dt = w[3 + c] - wrench_ref[3 + c];
                qf = qf + df * df;
                qt = qt + dt * dt;
            }
            acc = acc + (weights[0] * qf + weights[1] * qt);
        }
        cost[r] = acc;
    }
}

double ccmo_now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

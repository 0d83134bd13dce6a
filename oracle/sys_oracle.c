/*
 * sys_oracle.c -- CPU oracle for FloatingBaseSystemKinematics + ForwardEuler / FixedStepIntegrator
 * and for the J^T * wrench accumulation of FloatingBaseDynamicalSystem::dynamics.
 * TEST INFRASTRUCTURE ONLY; parity status in sys_oracle.h.
 *
 * Keeps the reference's expression structure (explicit matrix products, general 3x3 cofactor
 * inverse, one rounding per operation with -ffp-contract=off); the CUDA kernels use a different
 * derivation (w x column, symmetric inverse), so the two are independent.
 */
#include "sys_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "ccm_oracle.h"

/* ---- fixed-size algebra, row-major 3x3 ------------------------------------------------------ */

/* Eigen lazy product coefficient: ((a0*b0 + a1*b1) + a2*b2) */
static void mm3(const double a[9], const double b[9], double c[9])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double acc = a[3 * i] * b[j];
            acc = acc + a[3 * i + 1] * b[3 + j];
            acc = acc + a[3 * i + 2] * b[6 + j];
            c[3 * i + j] = acc;
        }
}

static void transpose3(const double a[9], double t[9])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) t[3 * i + j] = a[3 * j + i];
}

/* Eigen cofactor_3x3<i,j>: m(i1,j1)*m(i2,j2) - m(i1,j2)*m(i2,j1) */
static double cof3(const double m[9], int i, int j)
{
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[3 * i1 + j1] * m[3 * i2 + j2] - m[3 * i1 + j2] * m[3 * i2 + j1];
}

/* Eigen compute_inverse<Matrix3d, ., 3>: result(j,i) = cofactor<i,j> * invdet */
static void inverse3(const double m[9], double r[9])
{
    const double c00 = cof3(m, 0, 0), c10 = cof3(m, 1, 0), c20 = cof3(m, 2, 0);
    /* det = (cofactors_col0 .* matrix.col(0)).sum() */
    double det = c00 * m[0];
    det = det + c10 * m[3];
    det = det + c20 * m[6];
    const double invdet = 1.0 / det;
    r[0] = c00 * invdet;                 /* result.row(0) = cofactors_col0 * invdet */
    r[1] = c10 * invdet;
    r[2] = c20 * invdet;
    r[3] = cof3(m, 0, 1) * invdet;       /* result(1,0) */
    r[4] = cof3(m, 1, 1) * invdet;
    r[5] = cof3(m, 2, 1) * invdet;       /* result(1,2) */
    r[6] = cof3(m, 0, 2) * invdet;       /* result(2,0) */
    r[7] = cof3(m, 1, 2) * invdet;       /* result(2,1) */
    r[8] = cof3(m, 2, 2) * invdet;
}

/* ---- FloatingBaseSystemKinematics.cpp:36-73 ------------------------------------------------- */

void syso_kinematics_dynamics(double rho, const double twist[6], const double rot[9],
                              double pos_dot[3], double rot_dot[9])
{
    const double* w = twist + 3;
    /* baseLinearVelocity = baseTwist.head<3>()   :59 */
    for (int i = 0; i < 3; ++i) pos_dot[i] = twist[i];

    /* -baseRotation.colwise().cross(baseTwist.tail<3>())   :62 */
    double mcross[9];
    for (int j = 0; j < 3; ++j) {
        const double a0 = rot[j], a1 = rot[3 + j], a2 = rot[6 + j];
        mcross[j] = -(a1 * w[2] - a2 * w[1]);
        mcross[3 + j] = -(a2 * w[0] - a0 * w[2]);
        mcross[6 + j] = -(a0 * w[1] - a1 * w[0]);
    }
    /* m_rho / 2.0 * ((R * R^T).inverse() - I) * R   :63-66, parsed ((rho/2 * (inv - I)) * R) */
    double rt[9], rrt[9], inv[9], sM[9], baum[9];
    transpose3(rot, rt);
    mm3(rot, rt, rrt);
    inverse3(rrt, inv);
    const double half_rho = rho / 2.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            sM[3 * i + j] = half_rho * (inv[3 * i + j] - (i == j ? 1.0 : 0.0));
    mm3(sM, rot, baum);
    for (int i = 0; i < 9; ++i) rot_dot[i] = mcross[i] + baum[i];
}

/* ---- ForwardEuler.tpp:19-49, ForwardEuler.h:45-53 ------------------------------------------- */

void syso_forward_euler_step(double rho, double dT, const double twist[6], double pos[3],
                             double rot[9], int nj, const double* joint_vel, double* joint_pos)
{
    double pd[3], rd[9];
    syso_kinematics_dynamics(rho, twist, rot, pd, rd);
    /* std::get<I>(x) += std::get<I>(dx) * dT */
    for (int i = 0; i < 3; ++i) pos[i] = pos[i] + pd[i] * dT;
    for (int i = 0; i < 9; ++i) rot[i] = rot[i] + rd[i] * dT;
    for (int i = 0; i < nj; ++i) joint_pos[i] = joint_pos[i] + joint_vel[i] * dT;
}

/* ---- FixedStepIntegrator.tpp:19-76 ---------------------------------------------------------- */

int syso_integrate_schedule(double step_dT, double t0, double tf, double* dts, int cap)
{
    if (t0 > tf) return -1;          /* :32-38 */
    if (step_dT <= 0) return -1;     /* :40-46 */
    const int iterations = (int)ceil((tf - t0) / step_dT);   /* :48 */
    /* tf == t0 gives iterations == 0; the reference then compares `size_t i < iterations - 1`,
     * i.e. against SIZE_MAX, and never terminates.  Not reproducible: refused here. */
    if (iterations < 1) return -1;
    int count = 0;
    double current = t0;             /* :50 */
    for (int i = 0; i < iterations - 1; ++i) {                /* :51 (size_t i < int-1) */
        current = t0 + step_dT * i;  /* :53 */
        if (dts && count < cap) dts[count] = step_dT;
        ++count;
    }
    /* last step: dT = finalTime - currentTime   :64 */
    if (dts && count < cap) dts[count] = tf - current;
    ++count;
    return count;
}

int syso_integrate(double rho, double step_dT, double t0, double tf, const double twist[6],
                   double pos[3], double rot[9], int nj, const double* joint_vel,
                   double* joint_pos)
{
    const int count = syso_integrate_schedule(step_dT, t0, tf, NULL, 0);
    if (count < 0) return -1;
    double current = t0;
    for (int i = 0; i < count - 1; ++i) {
        current = t0 + step_dT * i;
        syso_forward_euler_step(rho, step_dT, twist, pos, rot, nj, joint_vel, joint_pos);
    }
    syso_forward_euler_step(rho, tf - current, twist, pos, rot, nj, joint_vel, joint_pos);
    return count;
}

/* ---- threads -------------------------------------------------------------------------------- */

typedef struct job job_t;
typedef void (*range_fn)(const job_t*, size_t begin, size_t end);

struct job {
    range_fn fn;
    size_t begin, end;
    /* shared */
    size_t n;
    double rho, dT;
    const double* const* twist_planes;
    double* const* pos_planes;
    double* const* rot_planes;
    /* rollout */
    int feet, horizon;
    const double* const* null_planes;
    const double* const* param_planes;
    const double* uniform;
    unsigned mask;
    double* const* wrench_planes;
    double* const* autodyn_planes;
    double* ctrl;
    const double* wrench_ref;
    const double* weights;
    double* chain_cost;
    /* generalized force */
    int cps, ncols;
    const double* const* in_planes;
    const double* jacobians;
    const double* base;
    double* out;
    int base_negate;
    /* floating-base Euler step */
    double* nu;
    double* joint_pos;
    double* base_pos;
    double* base_rot;
    /* mass-matrix solve */
    const double* mass;
    const double* reg;
    const double* known;
    const double* tau;
    double* acc;
};

static void* trampoline(void* arg)
{
    const job_t* j = (const job_t*)arg;
    j->fn(j, j->begin, j->end);
    return NULL;
}

static void parallel_for(job_t* proto, size_t n, int nthreads)
{
    if (nthreads <= 1 || n < (size_t)nthreads) {
        proto->fn(proto, 0, n);
        return;
    }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = *proto;
        jobs[t].begin = n * (size_t)t / (size_t)nthreads;
        jobs[t].end = n * (size_t)(t + 1) / (size_t)nthreads;
        pthread_create(&th[t], NULL, trampoline, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(jobs);
    free(th);
}

/* ---- batched Euler step --------------------------------------------------------------------- */

static void euler_range(const job_t* j, size_t begin, size_t end)
{
    for (size_t i = begin; i < end; ++i) {
        double tw[6], p[3], r[9];
        for (int c = 0; c < 6; ++c) tw[c] = j->twist_planes[c][i];
        for (int c = 0; c < 3; ++c) p[c] = j->pos_planes[c][i];
        for (int c = 0; c < 9; ++c) r[c] = j->rot_planes[c][i];
        syso_forward_euler_step(j->rho, j->dT, tw, p, r, 0, NULL, NULL);
        for (int c = 0; c < 3; ++c) j->pos_planes[c][i] = p[c];
        for (int c = 0; c < 9; ++c) j->rot_planes[c][i] = r[c];
    }
}

void syso_euler_step_batch_soa(size_t n, double rho, double dT, const double* const* twist_planes,
                               double* const* pos_planes, double* const* rot_planes, int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.fn = euler_range;
    j.rho = rho;
    j.dT = dT;
    j.twist_planes = twist_planes;
    j.pos_planes = pos_planes;
    j.rot_planes = rot_planes;
    parallel_for(&j, n, nthreads);
}

/* ---- fused rollout -------------------------------------------------------------------------- */

static void rollout_range(const job_t* j, size_t begin, size_t end)
{
    const size_t chains = j->n;
    ccmo_model m;
    ccmo_construct(&m);
    if (j->uniform)
        ccmo_initialize(&m, j->uniform[0], j->uniform[1], j->uniform[2], j->uniform[3]);
    for (size_t c = begin; c < end; ++c) {
        ccmo_transform tf, nf;
        for (int k = 0; k < 3; ++k) tf.pos[k] = j->pos_planes[k][c];
        for (int k = 0; k < 9; ++k) tf.rot[k] = j->rot_planes[k][c];
        for (int k = 0; k < 3; ++k) nf.pos[k] = j->null_planes[k][c];
        for (int k = 0; k < 9; ++k) nf.rot[k] = j->null_planes[3 + k] ? j->null_planes[3 + k][c] : 0.0;
        if (j->param_planes)
            ccmo_initialize(&m, j->param_planes[0][c], j->param_planes[1][c], j->param_planes[2][c],
                            j->param_planes[3][c]);
        double acc = 0.0;
        for (int t = 0; t < j->horizon; ++t) {
            const size_t i = (size_t)t * chains + c;
            ccmo_twist tw;
            for (int k = 0; k < 3; ++k) tw.lin[k] = j->twist_planes[k][i];
            for (int k = 0; k < 3; ++k) tw.ang[k] = j->twist_planes[3 + k][i];
            ccmo_set_state(&m, &tw, &tf);
            ccmo_set_null_force_transform(&m, &nf);
            const double* w = ccmo_get_contact_wrench(&m);
            if ((j->mask & CCMO_WRENCH) && j->wrench_planes)
                for (int k = 0; k < 6; ++k) j->wrench_planes[k][i] = w[k];
            if ((j->mask & CCMO_AUTODYN) && j->autodyn_planes) {
                const double* a = ccmo_get_autonomous_dynamics(&m);
                for (int k = 0; k < 6; ++k) j->autodyn_planes[k][i] = a[k];
            }
            if ((j->mask & CCMO_CTRL) && j->ctrl)
                memcpy(j->ctrl + 36 * i, ccmo_get_control_matrix(&m), 36 * sizeof(double));
            if (j->wrench_ref) {
                double qf = 0.0, qt = 0.0;
                for (int k = 0; k < 3; ++k) {
                    const double df = w[k] - j->wrench_ref[k];
                    const double dt = w[3 + k] - j->wrench_ref[3 + k];
                    qf = qf + df * df;
                    qt = qt + dt * dt;
                }
                acc = acc + (j->weights[0] * qf + j->weights[1] * qt);
            }
            double twv[6];
            memcpy(twv, &tw, sizeof(twv));
            syso_forward_euler_step(j->rho, j->dT, twv, tf.pos, tf.rot, 0, NULL, NULL);
        }
        for (int k = 0; k < 3; ++k) j->pos_planes[k][c] = tf.pos[k];
        for (int k = 0; k < 9; ++k) j->rot_planes[k][c] = tf.rot[k];
        if (j->chain_cost) j->chain_cost[c] = acc;
    }
}

void syso_rollout(size_t n_rollouts, int feet, int horizon, double dT, double rho,
                  const double* const* twist_planes, double* const* pos_planes,
                  double* const* rot_planes, const double* const* null_planes,
                  const double* const* param_planes, const double uniform[4], unsigned mask,
                  double* const* wrench_planes, double* const* autodyn_planes, double* ctrl,
                  const double wrench_ref[6], const double weights[2], double* chain_cost,
                  double* cost, int nthreads)
{
    const size_t chains = n_rollouts * (size_t)feet;
    double* cc = chain_cost;
    if (!cc && cost) cc = (double*)malloc(sizeof(double) * (chains ? chains : 1));
    job_t j;
    memset(&j, 0, sizeof(j));
    j.fn = rollout_range;
    j.n = chains;
    j.rho = rho;
    j.dT = dT;
    j.feet = feet;
    j.horizon = horizon;
    j.twist_planes = twist_planes;
    j.pos_planes = pos_planes;
    j.rot_planes = rot_planes;
    j.null_planes = null_planes;
    j.param_planes = param_planes;
    j.uniform = param_planes ? NULL : uniform;
    j.mask = mask;
    j.wrench_planes = wrench_planes;
    j.autodyn_planes = autodyn_planes;
    j.ctrl = ctrl;
    j.wrench_ref = wrench_ref;
    j.weights = weights;
    j.chain_cost = cc;
    parallel_for(&j, chains, nthreads);
    if (cost) {
        for (size_t r = 0; r < n_rollouts; ++r) {
            double acc = 0.0;
            for (int f = 0; f < feet; ++f) acc = acc + cc[r * (size_t)feet + (size_t)f];
            cost[r] = acc;
        }
    }
    if (cc != chain_cost) free(cc);
}

/* ---- J^T * wrench accumulation -------------------------------------------------------------- */

static void genforce_range(const job_t* j, size_t begin, size_t end)
{
    ccmo_model m;
    ccmo_construct(&m);
    if (j->uniform)
        ccmo_initialize(&m, j->uniform[0], j->uniform[1], j->uniform[2], j->uniform[3]);
    const int nc = j->ncols;
    for (size_t s = begin; s < end; ++s) {
        double* o = j->out + s * (size_t)nc;
        /* m_knownCoefficent starts from the bias terms (:191-196) */
        for (int q = 0; q < nc; ++q) o[q] = j->base ? j->base[s * (size_t)nc + (size_t)q] : 0.0;
        if (j->base_negate)   /* head = -baseWrench, tail = -jointTorques of the bias forces */
            for (int q = 0; q < nc; ++q) o[q] = -o[q];
        for (int c = 0; c < j->cps; ++c) {               /* for (contactWrench : contactWrenches) */
            const size_t i = s * (size_t)j->cps + (size_t)c;
            const double* const* P = j->in_planes;
            ccmo_twist tw;
            ccmo_transform tf, nf;
            for (int k = 0; k < 3; ++k) tw.lin[k] = P[k][i];
            for (int k = 0; k < 3; ++k) tw.ang[k] = P[3 + k][i];
            for (int k = 0; k < 3; ++k) tf.pos[k] = P[6 + k][i];
            for (int k = 0; k < 9; ++k) tf.rot[k] = P[9 + k] ? P[9 + k][i] : 0.0;
            for (int k = 0; k < 3; ++k) nf.pos[k] = P[18 + k][i];
            for (int k = 0; k < 9; ++k) nf.rot[k] = P[21 + k] ? P[21 + k][i] : 0.0;
            if (j->param_planes)
                ccmo_initialize(&m, j->param_planes[0][i], j->param_planes[1][i],
                                j->param_planes[2][i], j->param_planes[3][i]);
            ccmo_set_state(&m, &tw, &tf);            /* :221-222 */
            ccmo_set_null_force_transform(&m, &nf);
            const double* w = ccmo_get_contact_wrench(&m);
            if (j->wrench_planes)
                for (int k = 0; k < 6; ++k) j->wrench_planes[k][i] = w[k];
            /* m_knownCoefficent += J^T * wrench   :224-225 (product evaluated, then added) */
            const double* J = j->jacobians + i * 6u * (size_t)nc;
            for (int q = 0; q < nc; ++q) {
                double acc = J[q] * w[0];
                for (int r = 1; r < 6; ++r) acc = acc + J[(size_t)r * (size_t)nc + (size_t)q] * w[r];
                o[q] = o[q] + acc;
            }
        }
    }
}

void syso_generalized_force(size_t n_systems, int contacts_per_system, int ncols,
                            const double* const* in_planes, const double* const* param_planes,
                            const double uniform[4], const double* jacobians, const double* base,
                            double* out, double* const* wrench_planes, int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.fn = genforce_range;
    j.cps = contacts_per_system;
    j.ncols = ncols;
    j.in_planes = in_planes;
    j.param_planes = param_planes;
    j.uniform = param_planes ? NULL : uniform;
    j.jacobians = jacobians;
    j.base = base;
    j.out = out;
    j.wrench_planes = wrench_planes;
    parallel_for(&j, n_systems, nthreads);
}

/* ---- (M + reg).llt().solve(known) ------------------------------------------------------------ */

void syso_llt_solve_one(int nc, const double* mass, const double* reg, const double* rhs,
                        double* x, double* work)
{
    /* work = the matrix LLT factorises: M, or the evaluated sum M + reg (:236-239) */
    double* L = work;
    for (int i = 0; i < nc * nc; ++i) L[i] = reg ? mass[i] + reg[i] : mass[i];
    /* lower Cholesky, column by column; only the lower triangle is read */
    for (int j = 0; j < nc; ++j) {
        double d = L[j * nc + j];
        for (int k = 0; k < j; ++k) d = d - L[j * nc + k] * L[j * nc + k];
        const double ljj = sqrt(d);
        L[j * nc + j] = ljj;
        for (int i = j + 1; i < nc; ++i) {
            double t = L[i * nc + j];
            for (int k = 0; k < j; ++k) t = t - L[i * nc + k] * L[j * nc + k];
            L[i * nc + j] = t / ljj;
        }
    }
    for (int i = 0; i < nc; ++i) {                 /* L y = b */
        double t = rhs[i];
        for (int k = 0; k < i; ++k) t = t - L[i * nc + k] * x[k];
        x[i] = t / L[i * nc + i];
    }
    for (int i = nc - 1; i >= 0; --i) {            /* L^T x = y */
        double t = x[i];
        for (int k = i + 1; k < nc; ++k) t = t - L[k * nc + i] * x[k];
        x[i] = t / L[i * nc + i];
    }
}

static void llt_range(const job_t* j, size_t begin, size_t end)
{
    const int nc = j->ncols;
    double* work = (double*)malloc(sizeof(double) * ((size_t)nc * (size_t)nc + 2u * (size_t)nc));
    double* rhs = work + (size_t)nc * (size_t)nc;
    double* x = rhs + nc;
    for (size_t s = begin; s < end; ++s) {
        for (int q = 0; q < nc; ++q) rhs[q] = j->known[s * (size_t)nc + (size_t)q];
        if (j->tau)   /* m_knownCoefficent.tail(m_actuatedDoFs) += jointTorques  (:226-227) */
            for (int q = 6; q < nc; ++q) rhs[q] = rhs[q] + j->tau[s * (size_t)(nc - 6) + (size_t)(q - 6)];
        syso_llt_solve_one(nc, j->mass + s * (size_t)nc * (size_t)nc, j->reg, rhs, x, work);
        for (int q = 0; q < nc; ++q) j->acc[s * (size_t)nc + (size_t)q] = x[q];
    }
    free(work);
}

void syso_mass_matrix_solve(size_t n_systems, int ncols, const double* mass, const double* reg,
                            const double* known, const double* joint_torques, double* acc,
                            int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.fn = llt_range;
    j.ncols = ncols;
    j.mass = mass;
    j.reg = reg;
    j.known = known;
    j.tau = joint_torques;
    j.acc = acc;
    parallel_for(&j, n_systems, nthreads);
}

void syso_floating_base_acceleration(size_t n_systems, int contacts_per_system, int ncols,
                                     const double* const* in_planes,
                                     const double* const* param_planes, const double uniform[4],
                                     const double* jacobians, const double* bias_forces,
                                     const double* joint_torques, const double* mass,
                                     const double* reg, double* acc, double* const* wrench_planes,
                                     int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.fn = genforce_range;
    j.cps = contacts_per_system;
    j.ncols = ncols;
    j.in_planes = in_planes;
    j.param_planes = param_planes;
    j.uniform = param_planes ? NULL : uniform;
    j.jacobians = jacobians;
    j.base = bias_forces;
    j.base_negate = 1;
    j.out = acc;
    j.wrench_planes = wrench_planes;
    parallel_for(&j, n_systems, nthreads);
    syso_mass_matrix_solve(n_systems, ncols, mass, reg, acc, joint_torques, acc, nthreads);
}

/* ---- ForwardEuler<FloatingBaseDynamicalSystem>: x += dx * dT over the state tuple ---------------- */

static void fbd_euler_range(const job_t* j, size_t begin, size_t end)
{
    const int nc = j->ncols, nj = nc - 6;
    for (size_t s = begin; s < end; ++s) {
        double* nu = j->nu + s * (size_t)nc;
        /* base position, base rotation, joint positions: derivatives at the state before the step
         * (baseLinearVelocity = nu.head<3>(), the rotation rate of :139-145, jointVelocity) */
        syso_forward_euler_step(j->rho, j->dT, nu, j->base_pos + 3 * s, j->base_rot + 9 * s, nj > 0 ? nj : 0,
                                nj > 0 ? nu + 6 : NULL, nj > 0 ? j->joint_pos + s * (size_t)nj : NULL);
        /* base and joint velocity += acceleration * dT */
        for (int q = 0; q < nc; ++q) nu[q] = nu[q] + j->acc[s * (size_t)nc + (size_t)q] * j->dT;
    }
}

void syso_floating_base_euler_step(size_t n_systems, int ncols, double rho, double dT, const double* acc,
                                   double* nu, double* joint_pos, double* base_pos, double* base_rot,
                                   int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.fn = fbd_euler_range;
    j.ncols = ncols;
    j.rho = rho;
    j.dT = dT;
    j.acc = (double*)acc;
    j.nu = nu;
    j.joint_pos = joint_pos;
    j.base_pos = base_pos;
    j.base_rot = base_rot;
    parallel_for(&j, n_systems, nthreads);
}

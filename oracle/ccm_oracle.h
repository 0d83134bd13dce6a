/*
 * ccm_oracle.h -- CPU oracle for ContactModels::ContinuousContactModel.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path may include, link or call this.
 * Allowed users: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference leg.
 *
 * PARITY STATUS: pinned against THE REFERENCE'S OWN SOURCES, not against a binary linked with the
 * real Eigen/iDynTree.  Eigen, iDynTree and Catch2 are absent from this image (no network), and the
 * reference's tests hold no golden vectors for this path (ContinousContactModelTest.cpp has only
 * self-consistency checks with an unseeded random twist).  This file is a plain-C restatement of
 * the reference's closed-form algebra; it is pinned by
 *   (o)  oracle/_ref: ContinuousContactModel.cpp / ContactModel.cpp / StdImplementation.cpp
 *        compiled UNMODIFIED from /root/reference against stand-in Eigen/iDynTree headers
 *        (oracle/refbuild/README.md says exactly what the stand-ins are); this restatement must
 *        agree with that build BIT FOR BIT (tests/test_reference_build.py), and the reference's own
 *        test file passes on that build;
 *   (i)  exact rational evaluation of the same formulas (oracle/exact_golden.py ->
 *        tests/golden/ccm_exact_golden.npz), and
 *   (ii) the reference's three test properties restated in tests/test_oracle.py.
 * What remains unpinned: Eigen's internal evaluation order (an O(eps) effect, far below 1e-12).
 *
 * Follows (paths relative to /root/reference):
 *   src/ContactModels/src/ContinuousContactModel.cpp:16-274   all arithmetic
 *   src/ContactModels/src/ContactModel.cpp:12-92               lazy-flag protocol
 *   src/ContactModels/include/BipedalLocomotion/ContactModels/ContinuousContactModel.h:44-57  defaults
 *
 * Third-party semantics relied on (iDynTree >= 0.11.105, CI pin v1.1.0; Eigen >= 3.2.92; neither
 * is vendored under /root/reference):
 *   iDynTree::skew(v) = [[0,-v2,v1],[v2,0,-v0],[-v1,v0,0]]
 *   iDynTree::Twist / Wrench = linear(3) then angular(3)
 *   iDynTree::Transform     = Position(3) then Rotation(3x3 row-major)    (96 bytes)
 *   iDynTree::Matrix6x6 / MatrixDynSize = row-major
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, i.e. no FMA contraction, so every
 * product and sum is rounded once, as the reference's Release build on x86-64 without -mfma is).
 */
#ifndef CCM_ORACLE_H
#define CCM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double lin[3]; double ang[3]; } ccmo_twist;      /* iDynTree::Twist      */
typedef struct { double pos[3]; double rot[9]; } ccmo_transform;  /* iDynTree::Transform  */

/* One reference object (ContactModel base + ContinuousContactModel members). */
typedef struct {
    /* ContactModel.h:35-40 -- four lazy flags */
    int is_wrench_computed;
    int is_autodyn_computed;
    int is_ctrl_computed;
    int is_regressor_computed;
    /* ContactModel.h:43-56 -- result storage */
    double wrench[6];      /* force xyz, torque xyz */
    double autodyn[6];
    double ctrl[36];       /* row-major 6x6 */
    double regressor[12];  /* row-major 6x2 */
    /* ContinuousContactModel.h:44-57 */
    ccmo_transform frame;
    ccmo_transform null_force;
    ccmo_twist twist;
    double spring, damper, length, width;
} ccmo_model;

enum {
    CCMO_WRENCH = 1,
    CCMO_AUTODYN = 2,
    CCMO_CTRL = 4,
    CCMO_REGRESSOR = 8
};

/* ContinuousContactModel.cpp:16-22 (+ header defaults :44-57) */
void ccmo_construct(ccmo_model* m);
/* ContactModel.cpp:12-21 + ContinuousContactModel.cpp:24-65 (the four doubles the handler yields) */
void ccmo_initialize(ccmo_model* m, double length, double width, double spring, double damper);
/* ContactModel.cpp:35-48, ContinuousContactModel.cpp:72-77 */
void ccmo_set_state(ccmo_model* m, const ccmo_twist* twist, const ccmo_transform* transform);
/* ContactModel.cpp:23-33, ContinuousContactModel.cpp:67-70 */
void ccmo_set_null_force_transform(ccmo_model* m, const ccmo_transform* transform);
/* ContactModel.cpp:50-92 -- lazy getters, return pointers to internal storage */
const double* ccmo_get_contact_wrench(ccmo_model* m);
const double* ccmo_get_autonomous_dynamics(ccmo_model* m);
const double* ccmo_get_control_matrix(ccmo_model* m);
const double* ccmo_get_regressor(ccmo_model* m);
/* ContinuousContactModel.cpp:173-221 */
void ccmo_get_force_at_point(ccmo_model* m, double x, double y, double out[3]);
void ccmo_get_torque_generated_at_point(ccmo_model* m, double x, double y, double out[3]);

/*
 * Batch driver = the faithful per-instance path, one model object per thread:
 * per state setState, setNullForceTransform, then the getters selected by mask, results copied
 * out (how src/System/src/FloatingBaseSystemDynamics.cpp:221-225 drives the model).
 *
 * AoS: twists n*6, poses n*12, null_poses n*12; params = NULL (uniform, from `uniform[4]` =
 * length,width,spring,damper) or n*4 in that order.  Outputs (NULL if not in mask):
 * wrench n*6, autodyn n*6, ctrl n*36, regressor n*12.  nthreads <= 1 runs inline.
 */
void ccmo_eval_batch_aos(size_t n, const double* twists, const double* poses,
                         const double* null_poses, const double* params, const double uniform[4],
                         unsigned mask, double* wrench, double* autodyn, double* ctrl,
                         double* regressor, int nthreads);

/*
 * SoA flavour.  in_planes[30]: v(0-2) w(3-5) p(6-8) R row-major(9-17) p0(18-20) R0 row-major
 * (21-29); each plane n doubles.  param_planes[4] = length,width,spring,damper planes or NULL.
 * wrench_planes[6], autodyn_planes[6], regressor_planes[12] are plane pointers; ctrl is dense
 * n*36 row-major.
 */
void ccmo_eval_batch_soa(size_t n, const double* const* in_planes,
                         const double* const* param_planes, const double uniform[4],
                         unsigned mask, double* const* wrench_planes,
                         double* const* autodyn_planes, double* ctrl,
                         double* const* regressor_planes, int nthreads);

/*
 * Per-rollout cost used by the sampling-MPC path (no reference equivalent; definition in
 * DESIGN.md): cost[r] = sum over the rollout's evals e (in index order) of
 *   wf * |force_e - ref_force|^2 + wt * |torque_e - ref_torque|^2 ,  wrench_ref[6], weights[2].
 * wrench is AoS n*6 with n = n_rollouts*rollout_len.
 */
void ccmo_rollout_cost(size_t n_rollouts, size_t rollout_len, const double* wrench,
                       const double wrench_ref[6], const double weights[2], double* cost);

/* steady-clock seconds, for the CPU baseline timing */
double ccmo_now(void);

#ifdef __cplusplus
}
#endif
#endif /* CCM_ORACLE_H */

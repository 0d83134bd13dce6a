/*
 * sys_oracle.h -- CPU oracle for the two steps either side of the contact model in the reference's
 * System component.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path may include, link or call this.
 * Allowed users: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference leg.
 *
 * PARITY STATUS: kinematics, Euler step, integrate schedule and rollout are pinned against the
 * reference's own FloatingBaseSystemKinematics.cpp + ForwardEuler/FixedStepIntegrator templates
 * compiled unmodified into oracle/_ref against stand-in Eigen/iDynTree headers (bit-for-bit agreement
 * required, tests/test_reference_build.py; the reference's IntegratorTest.cpp passes on that build;
 * the 3x3 inverse() underneath is the stand-in's cofactor formula).  syso_generalized_force is
 * pinned the same way against FloatingBaseSystemDynamics.cpp compiled unmodified and run over a
 * KinDynComputations TEST DOUBLE (identity mass matrix; Jacobians, bias forces and frame states
 * injected -- iDynTree's rigid-body algorithms are not involved and not claimed).
 * The reference's only test of these functions (src/System/tests/IntegratorTest.cpp:80-126) compares
 * against a closed-form solution with tolerance 1e-3 on an unseeded random twist -- no golden
 * vectors.  This file restates the reference's arithmetic in plain C; it is also pinned by
 *   (i)  exact-rational single steps and 80-digit multi-step rollouts
 *        (oracle/exact_golden_sys.py -> tests/golden/sys_exact_golden.npz), and
 *   (ii) the reference test's property (rotation follows the axis-angle closed form within 1e-3,
 *        position is exactly linear) restated in tests/test_sys_oracle.py.
 *
 * Follows (paths relative to /root/reference):
 *   src/System/src/FloatingBaseSystemKinematics.cpp:36-73             dynamics()
 *   src/System/include/BipedalLocomotion/System/ForwardEuler.h:45-53   addArea: x += dx * dT
 *   src/System/include/BipedalLocomotion/System/ForwardEuler.tpp:19-49 oneStepIntegration
 *   src/System/include/BipedalLocomotion/System/FixedStepIntegrator.tpp:19-76  integrate()
 *   src/System/src/FloatingBaseSystemDynamics.cpp:199-226              known += J^T * wrench
 *   src/System/src/FloatingBaseSystemDynamics.cpp:188-196, 226-248     -bias, += torques, llt().solve
 *
 * Third-party semantics relied on (Eigen >= 3.2.92, not vendored; [from memory]):
 *   a.cross(b) = (a1 b2 - a2 b1, a2 b0 - a0 b2, a0 b1 - a1 b0);  M.colwise().cross(w) applies it
 *   to every column;  fixed-size 3x3 products sum k = 0,1,2 in order;  Matrix3d::inverse() is the
 *   cofactor formula (Eigen/src/LU/InverseImpl.h, compute_inverse<.,.,3>): cofactors of column 0,
 *   det = sum(cofactors_col0 .* col(0)), invdet = 1/det, every cofactor multiplied by invdet.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off).
 */
#ifndef SYS_ORACLE_H
#define SYS_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* FloatingBaseSystemKinematics::dynamics, base part (:58-68): pos_dot = twist.head<3>();
 * rot_dot = -R.colwise().cross(w) + rho/2 * ((R R^T)^-1 - I) * R.   rot is ROW-major 3x3. */
void syso_kinematics_dynamics(double rho, const double twist[6], const double rot[9],
                              double pos_dot[3], double rot_dot[9]);

/* ForwardEuler::oneStepIntegration on that system: dynamics, then x += dx * dT for position,
 * rotation and (optionally, nj > 0) the joint positions with dx = joint_vel (:70). */
void syso_forward_euler_step(double rho, double dT, const double twist[6], double pos[3],
                             double rot[9], int nj, const double* joint_vel, double* joint_pos);

/* FixedStepIntegrator::integrate(t0, tf) with sampling time step_dT and a constant control input
 * (:19-76).  Reproduces the reference's step schedule exactly, including its quirk: `currentTime`
 * is only advanced inside the loop, so for iterations >= 2 the final step is
 * tf - (t0 + step_dT*(iterations-2)).  Returns the number of Euler steps taken, or -1 where the
 * reference returns false (tf < t0, step_dT <= 0). */
int syso_integrate(double rho, double step_dT, double t0, double tf, const double twist[6],
                   double pos[3], double rot[9], int nj, const double* joint_vel,
                   double* joint_pos);

/* The step sizes integrate() uses (at most cap of them written); returns their count or -1. */
int syso_integrate_schedule(double step_dT, double t0, double tf, double* dts, int cap);

/* One Euler step for n independent systems, SoA planes of n doubles: twist_planes[6],
 * pos_planes[3] (in/out), rot_planes[9] row-major index (in/out). */
void syso_euler_step_batch_soa(size_t n, double rho, double dT, const double* const* twist_planes,
                               double* const* pos_planes, double* const* rot_planes, int nthreads);

/*
 * Fused sampling-MPC rollout (new; composition of the reference's pieces the way
 * FloatingBaseDynamicalSystem + ForwardEuler sequence them: evaluate the contact model at the
 * current state, then advance the state).  chains = n_rollouts * feet; chain c = rollout*feet+foot.
 * For t = 0..horizon-1, per chain: contact model <- (twist[t][c], pose_c), outputs per mask at
 * index t*chains + c (time-major); cost term; pose_c <- ForwardEuler step with twist[t][c].
 *   twist_planes[6]   each horizon*chains doubles, index t*chains + c
 *   pos_planes[3], rot_planes[9]       chains doubles each, initial pose in, final pose out
 *   null_planes[12]   null-force pose per chain (pos 0-2, rot row-major 3-11; rot third column
 *                     planes 5, 8, 11 may be NULL)
 *   param_planes[4]   per-chain length,width,spring,damper or NULL -> uniform[4]
 *   wrench_planes[6], autodyn_planes[6]: horizon*chains; ctrl: horizon*chains*36 dense
 *   chain_cost[c] = sum_t (wf|F-Fref|^2 + wt|T-Tref|^2) in t order;
 *   cost[r] = sum_foot chain_cost[r*feet+foot] in foot order.   Any output may be NULL.
 */
void syso_rollout(size_t n_rollouts, int feet, int horizon, double dT, double rho,
                  const double* const* twist_planes, double* const* pos_planes,
                  double* const* rot_planes, const double* const* null_planes,
                  const double* const* param_planes, const double uniform[4], unsigned mask,
                  double* const* wrench_planes, double* const* autodyn_planes, double* ctrl,
                  const double wrench_ref[6], const double weights[2], double* chain_cost,
                  double* cost, int nthreads);

/*
 * FloatingBaseDynamicalSystem::dynamics, contact part (:199-226): per system s,
 *   out[s] = base[s] + sum_{c in system, in order} J_c^T * wrench_c
 * with wrench_c = getContactWrench() of contact s*contacts_per_system + c evaluated from the SoA
 * state planes (as ccmo_eval_batch_soa), J_c the 6 x ncols row-major frame Jacobian
 * (iDynTree::MatrixDynSize).  jacobians: n_systems*contacts_per_system*6*ncols doubles;
 * base (may be NULL = zeros) and out: n_systems*ncols.  wrench_planes[6] optional (NULL).
 */
void syso_generalized_force(size_t n_systems, int contacts_per_system, int ncols,
                            const double* const* in_planes, const double* const* param_planes,
                            const double uniform[4], const double* jacobians, const double* base,
                            double* out, double* const* wrench_planes, int nthreads);

/*
 * FloatingBaseDynamicalSystem::dynamics, last step (:226-243):
 *   rhs = known;  rhs.tail(ncols - 6) += joint_torques;  acc = (mass [+ reg]).llt().solve(rhs)
 * Eigen's LLT (third party, >= 3.2.92, not vendored) restated from its published algorithm: the
 * lower Cholesky factor L of the LOWER triangle of A (column by column: d = A_jj - sum_k L_jk^2,
 * L_jj = sqrt(d), L_ij = (A_ij - sum_k L_ik L_jk) / L_jj, k ascending), then L y = b and L^T x = y by
 * substitution, one rounding per operation.  Eigen's own kernels sum the same products in a
 * vectorised / blocked order, so agreement with an Eigen binary is to rounding, scaled by the
 * conditioning of A -- not bit for bit; agreement with oracle/_ref (the reference's dynamics()
 * compiled over the stand-in Eigen, which evaluates LLT in exactly this order) IS bit for bit
 * (tests/test_reference_build.py).  A matrix that is not positive definite gives NaN here (sqrt of
 * a negative number); Eigen stops factorising at that column and solves with the partial factor.
 * mass: n_systems*ncols*ncols row-major; reg: ncols*ncols or NULL; known, acc: n_systems*ncols
 * (may alias); joint_torques: n_systems*(ncols-6) or NULL.
 */
void syso_mass_matrix_solve(size_t n_systems, int ncols, const double* mass, const double* reg,
                            const double* known, const double* joint_torques, double* acc,
                            int nthreads);

/* one system; work: ncols*ncols doubles of scratch */
void syso_llt_solve_one(int nc, const double* mass, const double* reg, const double* rhs, double* x,
                        double* work);

/* dynamics() from the bias forces on (:188-248): known = -bias_forces + sum_c J_c^T wrench_c
 * (syso_generalized_force with the negated bias as base), then syso_mass_matrix_solve in place. */
void syso_floating_base_acceleration(size_t n_systems, int contacts_per_system, int ncols,
                                     const double* const* in_planes,
                                     const double* const* param_planes, const double uniform[4],
                                     const double* jacobians, const double* bias_forces,
                                     const double* joint_torques, const double* mass,
                                     const double* reg, double* acc, double* const* wrench_planes,
                                     int nthreads);

/*
 * One ForwardEuler step of FloatingBaseDynamicalSystem (ForwardEuler.tpp:19-49: x = x0 + dx * dT over the
 * state tuple of FloatingBaseSystemDynamics.h:33-52), every derivative taken at the state BEFORE the
 * step: base position += nu.head<3>() * dT; base rotation += (rotation rate of
 * FloatingBaseSystemDynamics.cpp:139-145, the formula of FloatingBaseSystemKinematics) * dT;
 * joint positions += nu.tail * dT; nu += acc * dT.  acc as syso_floating_base_acceleration returns it.
 * nu n*ncols in/out, joint_pos n*(ncols-6) in/out (NULL when ncols == 6), base_pos n*3, base_rot n*9
 * row-major in/out.
 */
void syso_floating_base_euler_step(size_t n_systems, int ncols, double rho, double dT, const double* acc,
                                   double* nu, double* joint_pos, double* base_pos, double* base_rot,
                                   int nthreads);

#ifdef __cplusplus
}
#endif
#endif /* SYS_ORACLE_H */

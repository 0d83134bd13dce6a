// ref_driver.cpp -- C entry points that DRIVE THE REFERENCE'S OWN CLASSES, compiled from the
// reference's own sources under /root/reference (oracle/refbuild/Makefile) into
// oracle/_ref/libblf_reference.so.
//
// TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's CPU legs
// through oracle/ref_binding.py; never by the product.  No arithmetic of the path lives in this
// file: every number comes out of
//   BipedalLocomotion::ContactModels::ContinuousContactModel     (src/ContactModels/src/*.cpp)
//   BipedalLocomotion::ParametersHandler::StdImplementation      (src/ParametersHandler/src/*.cpp)
//   BipedalLocomotion::Estimators::RecursiveLeastSquare          (src/Estimators/src/*.cpp)
//   BipedalLocomotion::System::FloatingBaseSystemKinematics + ForwardEuler<> (src/System/...)
//   BipedalLocomotion::System::FloatingBaseDynamicalSystem (src/System/src/FloatingBaseSystemDynamics.cpp)
//     -- over a KinDynComputations TEST DOUBLE that returns injected Jacobians / bias forces / frame
//        states and an identity mass matrix (standin/idyntree/iDynTree/Model/StandinModel.h)
// called through their public interface, the way src/System/src/FloatingBaseSystemDynamics.cpp:211-225
// and the reference's tests call them.  Eigen and iDynTree are stand-ins (oracle/refbuild/standin).
#include <algorithm>
#include <chrono>
#include <cstddef>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include <iDynTree/Core/EigenHelpers.h>

#include <BipedalLocomotion/ContactModels/ContinuousContactModel.h>
#include <BipedalLocomotion/Estimators/RecursiveLeastSquare.h>
#include <BipedalLocomotion/ParametersHandler/StdImplementation.h>
#include <BipedalLocomotion/System/FloatingBaseSystemDynamics.h>
#include <BipedalLocomotion/System/FloatingBaseSystemKinematics.h>
#include <BipedalLocomotion/System/ForwardEuler.h>

using BipedalLocomotion::ContactModels::ContinuousContactModel;
using BipedalLocomotion::Estimators::RecursiveLeastSquare;
using BipedalLocomotion::ParametersHandler::IParametersHandler;
using BipedalLocomotion::ParametersHandler::StdImplementation;
using BipedalLocomotion::System::ContactWrench;
using BipedalLocomotion::System::FloatingBaseDynamicalSystem;
using BipedalLocomotion::System::FloatingBaseSystemKinematics;
using BipedalLocomotion::System::ForwardEuler;

namespace
{
enum : unsigned
{
    WRENCH = 1,
    AUTODYN = 2,
    CTRL = 4,
    REGRESSOR = 8
};

iDynTree::Transform makeTransform(const double* pose12)
{
    iDynTree::Transform t;
    t.setPosition(iDynTree::Position(pose12[0], pose12[1], pose12[2]));
    t.setRotation(iDynTree::Rotation(pose12 + 3, 3, 3)); // row-major 3x3
    return t;
}

iDynTree::Twist makeTwist(const double* tw6)
{
    iDynTree::Twist t;
    for (unsigned i = 0; i < 6; ++i)
        t(i) = tw6[i];
    return t;
}

bool initModel(ContinuousContactModel& model, const double par[4])
{
    std::shared_ptr<IParametersHandler> handler = std::make_shared<StdImplementation>();
    handler->setParameter("length", par[0]);
    handler->setParameter("width", par[1]);
    handler->setParameter("spring_coeff", par[2]);
    handler->setParameter("damper_coeff", par[3]);
    return model.initialize(handler);
}

// one model per thread, per state: setState, setNullForceTransform, the getters in `mask`
void evalRange(std::size_t begin, std::size_t end, const double* twists, const double* poses,
               const double* nullPoses, const double* params, const double uniform[4], unsigned mask,
               double* wrench, double* autodyn, double* ctrl, double* regressor, int* status)
{
    ContinuousContactModel model;
    if (params == nullptr && !initModel(model, uniform))
    {
        *status = -1;
        return;
    }
    for (std::size_t i = begin; i < end; ++i)
    {
        if (params != nullptr && !initModel(model, params + 4 * i))
        {
            *status = -1;
            return;
        }
        model.setState(makeTwist(twists + 6 * i), makeTransform(poses + 12 * i));
        model.setNullForceTransform(makeTransform(nullPoses + 12 * i));
        if (mask & WRENCH)
        {
            const iDynTree::Wrench& w = model.getContactWrench();
            for (unsigned k = 0; k < 6; ++k)
                wrench[6 * i + k] = w(k);
        }
        if (mask & AUTODYN)
        {
            const iDynTree::Vector6& f = model.getAutonomousDynamics();
            for (unsigned k = 0; k < 6; ++k)
                autodyn[6 * i + k] = f(k);
        }
        if (mask & CTRL)
        {
            const iDynTree::Matrix6x6& g = model.getControlMatrix();
            std::memcpy(ctrl + 36 * i, g.data(), 36 * sizeof(double));
        }
        if (mask & REGRESSOR)
        {
            const iDynTree::MatrixDynSize& y = model.getRegressor();
            std::memcpy(regressor + 12 * i, y.data(), 12 * sizeof(double));
        }
    }
}

template <class F> int runPartitioned(std::size_t n, int nthreads, F&& body)
{
    if (nthreads <= 1 || n < 2)
    {
        int status = 0;
        body(std::size_t(0), n, &status);
        return status;
    }
    const std::size_t nt = std::min<std::size_t>(std::size_t(nthreads), n);
    std::vector<std::thread> pool;
    std::vector<int> status(nt, 0);
    for (std::size_t t = 0; t < nt; ++t)
        pool.emplace_back(body, n * t / nt, n * (t + 1) / nt, &status[t]);
    for (auto& th : pool)
        th.join();
    for (int s : status)
        if (s != 0) return s;
    return 0;
}
} // namespace

extern "C" {

const char* blf_ref_description()
{
    return "reference sources compiled in place (ContactModels, ParametersHandler/StdImplementation, "
           "Estimators/RecursiveLeastSquare, System/FloatingBaseSystemKinematics + ForwardEuler, "
           "System/FloatingBaseSystemDynamics over a KinDynComputations test double) "
           "against stand-in Eigen/iDynTree headers";
}

double blf_ref_now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Same argument meaning as oracle/ccm_oracle.h::ccmo_eval_batch_aos.  Returns 0, or -1 if the
// reference's initialize() refused.
int blf_ref_ccm_eval_batch_aos(std::size_t n, const double* twists, const double* poses,
                               const double* null_poses, const double* params, const double uniform[4],
                               unsigned mask, double* wrench, double* autodyn, double* ctrl,
                               double* regressor, int nthreads)
{
    return runPartitioned(n, nthreads, [=](std::size_t b, std::size_t e, int* status) {
        evalRange(b, e, twists, poses, null_poses, params, uniform, mask, wrench, autodyn, ctrl, regressor,
                  status);
    });
}

// getForceAtPoint / getTorqueGeneratedAtPoint of ONE model state at npts surface points xy[npts*2].
int blf_ref_ccm_surface_points(const double twist[6], const double pose[12], const double null_pose[12],
                               const double par[4], std::size_t npts, const double* xy, double* force,
                               double* torque)
{
    ContinuousContactModel model;
    if (!initModel(model, par)) return -1;
    model.setState(makeTwist(twist), makeTransform(pose));
    model.setNullForceTransform(makeTransform(null_pose));
    for (std::size_t i = 0; i < npts; ++i)
    {
        if (force)
        {
            const iDynTree::Force f = model.getForceAtPoint(xy[2 * i], xy[2 * i + 1]);
            for (unsigned k = 0; k < 3; ++k)
                force[3 * i + k] = f(k);
        }
        if (torque)
        {
            const iDynTree::Torque t = model.getTorqueGeneratedAtPoint(xy[2 * i], xy[2 * i + 1]);
            for (unsigned k = 0; k < 3; ++k)
                torque[3 * i + k] = t(k);
        }
    }
    return 0;
}

// The lazy-cache protocol of ContactModel.cpp:12-92 observed from outside: after setState the four
// getters are called, then springCoeff() is overwritten through the mutable reference and the
// wrench is read again WITHOUT a setter in between (stale by design), then setState is repeated and
// the wrench read once more.  out[0..5] first wrench, out[6..11] stale read, out[12..17] fresh read.
int blf_ref_ccm_stale_cache_probe(const double twist[6], const double pose[12], const double null_pose[12],
                                  const double par[4], double new_spring, double out[18])
{
    ContinuousContactModel model;
    if (!initModel(model, par)) return -1;
    model.setState(makeTwist(twist), makeTransform(pose));
    model.setNullForceTransform(makeTransform(null_pose));
    for (unsigned k = 0; k < 6; ++k)
        out[k] = model.getContactWrench()(k);
    model.springCoeff() = new_spring;
    for (unsigned k = 0; k < 6; ++k)
        out[6 + k] = model.getContactWrench()(k);
    model.setState(makeTwist(twist), makeTransform(pose));
    for (unsigned k = 0; k < 6; ++k)
        out[12 + k] = model.getContactWrench()(k);
    return 0;
}

// initialize() error behaviour: which == 0 expired handler, 1..4 the key of that index missing
// (length, width, spring_coeff, damper_coeff), 5 a key present with the wrong type (int).
// Returns what initialize() returned (0 / 1).
int blf_ref_ccm_initialize_probe(int which)
{
    ContinuousContactModel model;
    if (which == 0)
    {
        std::weak_ptr<IParametersHandler> expired;
        {
            std::shared_ptr<IParametersHandler> tmp = std::make_shared<StdImplementation>();
            expired = tmp;
        }
        return model.initialize(expired) ? 1 : 0;
    }
    static const char* keys[4] = {"length", "width", "spring_coeff", "damper_coeff"};
    std::shared_ptr<IParametersHandler> handler = std::make_shared<StdImplementation>();
    for (int k = 0; k < 4; ++k)
    {
        if (which == k + 1) continue;
        if (which == 5 && k == 2)
            handler->setParameter(keys[k], 2000);
        else
            handler->setParameter(keys[k], 1.0);
    }
    return model.initialize(handler) ? 1 : 0;
}

// RecursiveLeastSquare: initialize from a StdImplementation handler, then nsteps times
// setMeasurements(z_s) + advance() with the regressor callback returning Y_s (m x p row-major).
// theta_out nsteps*p, P_out nsteps*p*p (row-major) after every step.  Returns 0, -1 (initialize),
// -2 (advance).
int blf_ref_rls_run(int p, int m, const double* measurement_cov, double lambda, const double* state0,
                    const double* state_cov_diag, int nsteps, const double* Y, const double* z,
                    double* theta_out, double* P_out)
{
    std::shared_ptr<IParametersHandler> handler = std::make_shared<StdImplementation>();
    handler->setParameter("measurement_covariance", std::vector<double>(measurement_cov, measurement_cov + m));
    handler->setParameter("lambda", lambda);
    handler->setParameter("state", std::vector<double>(state0, state0 + p));
    handler->setParameter("state_covariance", std::vector<double>(state_cov_diag, state_cov_diag + p));
    RecursiveLeastSquare estimator;
    if (!estimator.initialize(handler)) return -1;
    const double* current = Y;
    estimator.setRegressorFunction([&]() {
        iDynTree::MatrixDynSize reg(static_cast<std::size_t>(m), static_cast<std::size_t>(p));
        std::memcpy(reg.data(), current, sizeof(double) * std::size_t(m * p));
        return reg;
    });
    for (int s = 0; s < nsteps; ++s)
    {
        current = Y + std::size_t(s) * std::size_t(m * p);
        estimator.setMeasurements(iDynTree::VectorDynSize(z + std::size_t(s) * std::size_t(m), std::size_t(m)));
        if (!estimator.advance()) return -2;
        const iDynTree::VectorDynSize& th = estimator.parametersExpectedValue();
        const iDynTree::MatrixDynSize& P = estimator.parametersCovarianceMatrix();
        std::memcpy(theta_out + std::size_t(s) * std::size_t(p), th.data(), sizeof(double) * std::size_t(p));
        std::memcpy(P_out + std::size_t(s) * std::size_t(p * p), P.data(), sizeof(double) * std::size_t(p * p));
    }
    return 0;
}

namespace
{
struct KinematicsRig
{
    std::shared_ptr<FloatingBaseSystemKinematics> system;
    std::unique_ptr<ForwardEuler<FloatingBaseSystemKinematics>> integrator;
    bool ok = false;
    KinematicsRig(double rho, double stepDT)
    {
        system = std::make_shared<FloatingBaseSystemKinematics>();
        std::shared_ptr<IParametersHandler> handler = std::make_shared<StdImplementation>();
        handler->setParameter("rho", rho);
        ok = system->initalize(handler);
        integrator = std::make_unique<ForwardEuler<FloatingBaseSystemKinematics>>(stepDT);
        ok = ok && integrator->setDynamicalSystem(system);
    }
    bool set(const double twist[6], const double pos[3], const double rotRowMajor[9], int nj,
             const double* jointVel, const double* jointPos)
    {
        Eigen::Matrix<double, 6, 1> tw;
        Eigen::Vector3d p;
        Eigen::Matrix3d R;
        for (int i = 0; i < 6; ++i)
            tw(i) = twist[i];
        for (int i = 0; i < 3; ++i)
        {
            p(i) = pos[i];
            for (int j = 0; j < 3; ++j)
                R(i, j) = rotRowMajor[3 * i + j];
        }
        Eigen::VectorXd jv(nj), jp(nj);
        for (int i = 0; i < nj; ++i)
        {
            jv(i) = jointVel[i];
            jp(i) = jointPos[i];
        }
        return system->setControlInput({tw, jv}) && system->setState({p, R, jp});
    }
    bool setTwist(const double twist[6])
    {
        Eigen::Matrix<double, 6, 1> tw;
        for (int i = 0; i < 6; ++i)
            tw(i) = twist[i];
        return system->setControlInput({tw, Eigen::VectorXd(0)});
    }
    void get(double pos[3], double rotRowMajor[9], int nj, double* jointPos) const
    {
        const auto& [p, R, jp] = integrator->getSolution();
        for (int i = 0; i < 3; ++i)
        {
            pos[i] = p(i);
            for (int j = 0; j < 3; ++j)
                rotRowMajor[3 * i + j] = R(i, j);
        }
        for (int i = 0; i < nj; ++i)
            jointPos[i] = jp(i);
    }
};
} // namespace

// FloatingBaseSystemKinematics::dynamics at one state: pos_dot[3], rot_dot[9] row-major.
int blf_ref_kin_dynamics(double rho, const double twist[6], const double rot[9], double pos_dot[3],
                         double rot_dot[9])
{
    KinematicsRig rig(rho, 1.0);
    const double zero[3] = {0, 0, 0};
    if (!rig.ok || !rig.set(twist, zero, rot, 0, nullptr, nullptr)) return -1;
    FloatingBaseSystemKinematics::StateDerivativeType dx;
    if (!rig.system->dynamics(0.0, dx)) return -2;
    for (int i = 0; i < 3; ++i)
    {
        pos_dot[i] = std::get<0>(dx)(i);
        for (int j = 0; j < 3; ++j)
            rot_dot[3 * i + j] = std::get<1>(dx)(i, j);
    }
    return 0;
}

// ForwardEuler<FloatingBaseSystemKinematics>(step_dT).integrate(t0, tf) with a constant control
// input; pos / rot (row-major) / joint_pos are updated in place.  Returns 0, -1 (setup), -2
// (integrate returned false).
int blf_ref_kin_integrate(double rho, double step_dT, double t0, double tf, const double twist[6],
                          double pos[3], double rot[9], int nj, const double* joint_vel, double* joint_pos)
{
    KinematicsRig rig(rho, step_dT);
    if (!rig.ok || !rig.set(twist, pos, rot, nj, joint_vel, joint_pos)) return -1;
    if (!rig.integrator->integrate(t0, tf)) return -2;
    rig.get(pos, rot, nj, joint_pos);
    return 0;
}

// The sampling-MPC rollout as a composition of the reference's objects (argument meaning as
// oracle/sys_oracle.h::syso_rollout, but AoS for brevity): per chain c and step t
//   model.setState(twist[t][c], pose_c); model.setNullForceTransform(null_c); read the getters in
//   `mask` to index t*chains + c; pose_c <- ForwardEuler(dT).integrate(0, dT) under twist[t][c].
// twists: horizon*chains*6; poses (in/out): chains*12; null_poses chains*12; params chains*4 or
// NULL -> uniform.  No cost here: the cost is this repository's addition, the tests apply the
// oracle's cost definition to these wrenches.
int blf_ref_rollout(std::size_t chains, int horizon, double dT, double rho, const double* twists,
                    double* poses, const double* null_poses, const double* params, const double uniform[4],
                    unsigned mask, double* wrench, double* autodyn, double* ctrl, int nthreads)
{
    return runPartitioned(chains, nthreads, [=](std::size_t b, std::size_t e, int* status) {
        ContinuousContactModel model;
        KinematicsRig rig(rho, dT);
        if (!rig.ok)
        {
            *status = -1;
            return;
        }
        for (std::size_t c = b; c < e; ++c)
        {
            if (!initModel(model, params ? params + 4 * c : uniform))
            {
                *status = -1;
                return;
            }
            double* pose = poses + 12 * c;
            const iDynTree::Transform nullT = makeTransform(null_poses + 12 * c);
            if (!rig.set(twists + 6 * c, pose, pose + 3, 0, nullptr, nullptr))
            {
                *status = -1;
                return;
            }
            for (int t = 0; t < horizon; ++t)
            {
                const std::size_t idx = std::size_t(t) * chains + c;
                const double* tw = twists + 6 * idx;
                model.setState(makeTwist(tw), makeTransform(pose));
                model.setNullForceTransform(nullT);
                if (mask & WRENCH)
                    for (unsigned k = 0; k < 6; ++k)
                        wrench[6 * idx + k] = model.getContactWrench()(k);
                if (mask & AUTODYN)
                    for (unsigned k = 0; k < 6; ++k)
                        autodyn[6 * idx + k] = model.getAutonomousDynamics()(k);
                if (mask & CTRL) std::memcpy(ctrl + 36 * idx, model.getControlMatrix().data(), 36 * sizeof(double));
                if (!rig.setTwist(tw) || !rig.integrator->integrate(0.0, dT))
                {
                    *status = -2;
                    return;
                }
                rig.get(pose, pose + 3, 0, nullptr);
            }
        }
    });
}

// State in / out of one ForwardEuler<FloatingBaseDynamicalSystem> step (all arrays per system, row-major):
// nu = [base velocity (6); joint velocity], joint positions, base position (3), base rotation (9).
struct EulerStepIO
{
    double rho, dT;
    const double* nu;
    const double* jointPos;
    const double* basePos;
    const double* baseRot;
    double* nuOut;
    double* jointPosOut;
    double* basePosOut;
    double* baseRotOut;
};

// FloatingBaseDynamicalSystem::dynamics, contact part (FloatingBaseSystemDynamics.cpp:188-229), run
// from the reference's own source: per system the KinDynComputations test double is loaded with
//   mass matrix = identity (so the final llt().solve() returns m_knownCoefficent unchanged),
//   generalized bias forces h = -base (so that -h = base),
//   per contact c (frame index c): Jacobian J_c (6 x ncols row-major), frame velocity, world transform,
// the joint torques are zero, and dynamics() is called; out = [baseAcceleration; jointAcceleration]
// = base + sum_c J_c^T * wrench_c in the reference's own order.  Argument meaning as
// oracle/sys_oracle.h::syso_generalized_force, AoS: twists n*6, poses n*12, null_poses n*12 with
// n = n_systems*contacts_per_system; jacobians n*6*ncols; base / out n_systems*ncols; wrench n*6 or NULL.
// The same routine, given mass matrices / joint torques / a regularisation term and the bias forces
// themselves (base_is_bias), runs the whole of dynamics() for blf_ref_floating_base_dynamics below.
static int dynamicsOverTestDouble(std::size_t n_systems, int contacts_per_system, int ncols, const double* twists,
                                  const double* poses, const double* null_poses, const double* params,
                                  const double uniform[4], const double* jacobians, const double* base,
                                  bool base_is_bias, const double* joint_torques, const double* mass_matrices,
                                  const double* regularization, double* out, double* wrench, int nthreads,
                                  const EulerStepIO* step = nullptr)
{
    if (ncols < 6 || contacts_per_system < 1) return -3;
    return runPartitioned(n_systems, nthreads, [=](std::size_t b, std::size_t e, int* status) {
        const std::size_t nc = std::size_t(ncols), dofs = nc - 6, cps = std::size_t(contacts_per_system);
        for (std::size_t s = b; s < e; ++s)
        {
            auto kinDyn = std::make_shared<iDynTree::KinDynComputations>();
            kinDyn->standinSetModel(dofs);
            iDynTree::MatrixDynSize mass(nc, nc);
            if (mass_matrices)
                std::memcpy(mass.data(), mass_matrices + s * nc * nc, sizeof(double) * nc * nc);
            else
                for (std::size_t i = 0; i < nc; ++i)
                    mass(i, i) = 1.0;
            kinDyn->standinSetMassMatrix(mass);
            iDynTree::FreeFloatingGeneralizedTorques h(kinDyn->model());
            for (std::size_t q = 0; q < nc; ++q)
            {
                const double v = base ? (base_is_bias ? base[s * nc + q] : -base[s * nc + q]) : -0.0;
                if (q < 6)
                    h.baseWrench()(static_cast<unsigned>(q)) = v;
                else
                    h.jointTorques()(q - 6) = v;
            }
            kinDyn->standinSetBiasForces(h);

            std::vector<std::shared_ptr<ContinuousContactModel>> models;
            std::vector<ContactWrench> contacts;
            for (std::size_t c = 0; c < cps; ++c)
            {
                const std::size_t i = s * cps + c;
                iDynTree::MatrixDynSize J(6, nc);
                std::memcpy(J.data(), jacobians + i * 6 * nc, sizeof(double) * 6 * nc);
                kinDyn->standinSetFrame(static_cast<iDynTree::FrameIndex>(c), J, makeTwist(twists + 6 * i),
                                        makeTransform(poses + 12 * i));
                auto model = std::make_shared<ContinuousContactModel>();
                if (!initModel(*model, params ? params + 4 * i : uniform))
                {
                    *status = -1;
                    return;
                }
                model->setNullForceTransform(makeTransform(null_poses + 12 * i));
                models.push_back(model);
                contacts.emplace_back(static_cast<iDynTree::FrameIndex>(c), model);
            }

            auto systemPtr = std::make_shared<FloatingBaseDynamicalSystem>();
            FloatingBaseDynamicalSystem& system = *systemPtr;
            if (!system.setKinDyn(kinDyn))
            {
                *status = -1;
                return;
            }
            if (step)
            {
                std::shared_ptr<IParametersHandler> handler = std::make_shared<StdImplementation>();
                handler->setParameter("rho", step->rho);
                if (!system.initalize(handler))
                {
                    *status = -1;
                    return;
                }
            }
            const Eigen::Index nd = Eigen::Index(dofs);
            Eigen::Matrix<double, 6, 1> baseVelocity;
            baseVelocity.setZero();
            Eigen::VectorXd zeros(nd);
            zeros.setZero();
            Eigen::Vector3d basePosition;
            basePosition.setZero();
            Eigen::Matrix3d baseOrientation;
            baseOrientation.setIdentity();
            Eigen::VectorXd torques(nd);
            torques.setZero();
            if (joint_torques)
                for (Eigen::Index q = 0; q < nd; ++q)
                    torques(q) = joint_torques[s * dofs + std::size_t(q)];
            if (regularization)
            {
                const Eigen::Index ncI = Eigen::Index(nc);
                Eigen::MatrixXd reg(ncI, ncI);
                for (std::size_t i = 0; i < nc; ++i)
                    for (std::size_t k = 0; k < nc; ++k)
                        reg(Eigen::Index(i), Eigen::Index(k)) = regularization[i * nc + k];
                if (!system.setMassMatrixRegularization(reg))
                {
                    *status = -1;
                    return;
                }
            }
            Eigen::VectorXd jointVelocity = zeros, jointPosition = zeros;
            if (step)
            {
                for (Eigen::Index q = 0; q < 6; ++q)
                    baseVelocity(q) = step->nu[s * nc + std::size_t(q)];
                for (Eigen::Index q = 0; q < nd; ++q)
                {
                    jointVelocity(q) = step->nu[s * nc + 6 + std::size_t(q)];
                    jointPosition(q) = step->jointPos[s * dofs + std::size_t(q)];
                }
                for (Eigen::Index i = 0; i < 3; ++i)
                {
                    basePosition(i) = step->basePos[s * 3 + std::size_t(i)];
                    for (Eigen::Index k = 0; k < 3; ++k)
                        baseOrientation(i, k) = step->baseRot[s * 9 + std::size_t(3 * i + k)];
                }
            }
            if (!system.setState({baseVelocity, jointVelocity, basePosition, baseOrientation, jointPosition})
                || !system.setControlInput({torques, contacts}))
            {
                *status = -1;
                return;
            }
            FloatingBaseDynamicalSystem::StateDerivativeType dx;
            if (!system.dynamics(0.0, dx))
            {
                *status = -2;
                return;
            }
            if (step)
            {   // x <- x + dx * dT through the reference's own integrator (ForwardEuler.tpp:19-49)
                ForwardEuler<FloatingBaseDynamicalSystem> integrator(step->dT);
                if (!integrator.setDynamicalSystem(systemPtr) || !integrator.integrate(0.0, step->dT))
                {
                    *status = -2;
                    return;
                }
                const auto& [v6, jv, p3, R3, jp] = integrator.getSolution();
                for (Eigen::Index q = 0; q < 6; ++q)
                    step->nuOut[s * nc + std::size_t(q)] = v6(q);
                for (Eigen::Index q = 0; q < nd; ++q)
                {
                    step->nuOut[s * nc + 6 + std::size_t(q)] = jv(q);
                    step->jointPosOut[s * dofs + std::size_t(q)] = jp(q);
                }
                for (Eigen::Index i = 0; i < 3; ++i)
                {
                    step->basePosOut[s * 3 + std::size_t(i)] = p3(i);
                    for (Eigen::Index k = 0; k < 3; ++k)
                        step->baseRotOut[s * 9 + std::size_t(3 * i + k)] = R3(i, k);
                }
            }
            for (std::size_t q = 0; q < 6; ++q)
                out[s * nc + q] = std::get<0>(dx)(Eigen::Index(q));
            for (std::size_t q = 6; q < nc; ++q)
                out[s * nc + q] = std::get<1>(dx)(Eigen::Index(q - 6));
            if (wrench)
                for (std::size_t c = 0; c < cps; ++c)
                    for (unsigned k = 0; k < 6; ++k)
                        wrench[6 * (s * cps + c) + k] = models[c]->getContactWrench()(k);
        }
    });
}

int blf_ref_generalized_force(std::size_t n_systems, int contacts_per_system, int ncols, const double* twists,
                              const double* poses, const double* null_poses, const double* params,
                              const double uniform[4], const double* jacobians, const double* base, double* out,
                              double* wrench, int nthreads)
{
    return dynamicsOverTestDouble(n_systems, contacts_per_system, ncols, twists, poses, null_poses, params,
                                  uniform, jacobians, base, false, nullptr, nullptr, nullptr, out, wrench,
                                  nthreads);
}

// FloatingBaseDynamicalSystem::dynamics from the bias forces on (FloatingBaseSystemDynamics.cpp:188-248),
// run from the reference's own source: the test double is loaded per system with the given mass matrix
// (ncols x ncols row-major), generalized bias forces h (ncols: base wrench, joint torques) and the
// contact frames as above; joint_torques (n_systems*(ncols-6)) or NULL = zeros; regularization
// (ncols x ncols row-major, through setMassMatrixRegularization) or NULL.
// out = [baseAcceleration; jointAcceleration] = (M + reg).llt().solve(-h + sum J^T wrench + [0; tau]).
int blf_ref_floating_base_dynamics(std::size_t n_systems, int contacts_per_system, int ncols, const double* twists,
                                   const double* poses, const double* null_poses, const double* params,
                                   const double uniform[4], const double* jacobians, const double* bias_forces,
                                   const double* joint_torques, const double* mass_matrices,
                                   const double* regularization, double* out, double* wrench, int nthreads)
{
    if (!bias_forces || !mass_matrices) return -3;
    return dynamicsOverTestDouble(n_systems, contacts_per_system, ncols, twists, poses, null_poses, params,
                                  uniform, jacobians, bias_forces, true, joint_torques, mass_matrices,
                                  regularization, out, wrench, nthreads);
}

// One ForwardEuler<FloatingBaseDynamicalSystem>(dT).integrate(0, dT) per system, from the reference's own
// sources: dynamics() as blf_ref_floating_base_dynamics (out = the acceleration at the state before the
// step), then x += dx * dT over the whole state tuple.  nu / nu_out n*ncols, joint_pos n*(ncols-6),
// base_pos n*3, base_rot n*9 row-major.
int blf_ref_floating_base_euler_step(std::size_t n_systems, int contacts_per_system, int ncols, const double* twists,
                                     const double* poses, const double* null_poses, const double* params,
                                     const double uniform[4], const double* jacobians, const double* bias_forces,
                                     const double* joint_torques, const double* mass_matrices,
                                     const double* regularization, double rho, double dT, const double* nu,
                                     const double* joint_pos, const double* base_pos, const double* base_rot,
                                     double* acc, double* nu_out, double* joint_pos_out, double* base_pos_out,
                                     double* base_rot_out, int nthreads)
{
    if (!bias_forces || !mass_matrices || !nu || !base_pos || !base_rot) return -3;
    const EulerStepIO io{rho, dT, nu, joint_pos, base_pos, base_rot, nu_out, joint_pos_out, base_pos_out, base_rot_out};
    return dynamicsOverTestDouble(n_systems, contacts_per_system, ncols, twists, poses, null_poses, params,
                                  uniform, jacobians, bias_forces, true, joint_torques, mass_matrices,
                                  regularization, acc, nullptr, nthreads, &io);
}

} // extern "C"

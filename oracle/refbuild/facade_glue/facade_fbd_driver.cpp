// facade_fbd_driver.cpp -- TEST INFRASTRUCTURE (see oracle/refbuild/README.md), not product code.
//
// A C interface over the PRODUCT's C++ facade classes System::FloatingBaseDynamicalSystem +
// System::ForwardEuler (bipedal_locomotion_framework_b200/cpp), with the argument list of
// blf_ref_floating_base_euler_step in ref_driver.cpp -- which runs the REFERENCE'S own
// FloatingBaseSystemDynamics.cpp + ForwardEuler.tpp over a KinDynComputations test double.  A test hands
// both the same arrays and compares: the facade class on the GPU against the reference class on the CPU,
// each driven through its public methods only (setKinDyn, setMassMatrixRegularization, setState,
// setControlInput, dynamics, integrate).  The KinDynComputations object here answers with the caller's
// arrays, like the reference build's test double.  No arithmetic of the path lives in this file.
#include <cstddef>
#include <map>
#include <memory>
#include <vector>

#include <BipedalLocomotion/ContactModels/ContinuousContactModel.h>
#include <BipedalLocomotion/ParametersHandler/StdImplementation.h>
#include <BipedalLocomotion/System/FloatingBaseSystemDynamics.h>
#include <BipedalLocomotion/System/ForwardEuler.h>

using namespace BipedalLocomotion::ContactModels;
using namespace BipedalLocomotion::ParametersHandler;
using namespace BipedalLocomotion::System;

namespace
{
class ArraysKinDyn : public iDynTree::KinDynComputations
{
public:
    struct Frame
    {
        iDynTree::MatrixDynSize jacobian;
        iDynTree::Twist velocity;
        iDynTree::Transform transform;
    };
    iDynTree::Model robot;
    iDynTree::MatrixDynSize mass;
    iDynTree::FreeFloatingGeneralizedTorques bias;
    std::map<iDynTree::FrameIndex, Frame> frames;

    explicit ArraysKinDyn(std::size_t dofs) : robot(dofs), bias(robot) {}
    const iDynTree::Model& model() const override { return robot; }
    bool setRobotState(const iDynTree::Transform&, const iDynTree::VectorDynSize& s, const iDynTree::Twist&,
                       const iDynTree::VectorDynSize& ds, const iDynTree::Vector3&) override
    {
        return s.size() == robot.getNrOfDOFs() && ds.size() == robot.getNrOfDOFs();
    }
    bool getFreeFloatingMassMatrix(iDynTree::MatrixDynSize& out) override
    {
        out = mass;
        return true;
    }
    bool generalizedBiasForces(iDynTree::FreeFloatingGeneralizedTorques& out) override
    {
        out = bias;
        return true;
    }
    bool getFrameFreeFloatingJacobian(const iDynTree::FrameIndex frame, iDynTree::MatrixDynSize& out) override
    {
        auto it = frames.find(frame);
        if (it == frames.end()) return false;
        out = it->second.jacobian;
        return true;
    }
    iDynTree::Twist getFrameVel(const iDynTree::FrameIndex frame) override { return frames.at(frame).velocity; }
    iDynTree::Transform getWorldTransform(const iDynTree::FrameIndex frame) override
    {
        return frames.at(frame).transform;
    }
};

iDynTree::Transform transformFromRow(const double* row) // position (3), rotation (9, row-major)
{
    iDynTree::Rotation R;
    for (int k = 0; k < 9; ++k) R.data()[k] = row[3 + k];
    return iDynTree::Transform(R, iDynTree::Position(row[0], row[1], row[2]));
}
} // namespace

extern "C" int blf_facade_floating_base_euler_step(
    std::size_t n_systems, int contacts_per_system, int ncols, const double* twists, const double* poses,
    const double* null_poses, const double* params, const double uniform[4], const double* jacobians,
    const double* bias_forces, const double* joint_torques, const double* mass_matrices, const double* regularization,
    double rho, double dT, const double* nu, const double* joint_pos, const double* base_pos, const double* base_rot,
    double* acc, double* nu_out, double* joint_pos_out, double* base_pos_out, double* base_rot_out)
{
    if (ncols < 6 || contacts_per_system < 0) return -1;
    const std::size_t n = static_cast<std::size_t>(ncols), dofs = n - 6, cps = static_cast<std::size_t>(contacts_per_system);
    auto kinDyn = std::make_shared<ArraysKinDyn>(dofs);
    kinDyn->mass.resize(n, n);
    std::vector<std::shared_ptr<ContinuousContactModel>> models;
    std::vector<ContactWrench> contacts;
    for (std::size_t c = 0; c < cps; ++c)
    {
        models.push_back(std::make_shared<ContinuousContactModel>());
        contacts.emplace_back(static_cast<iDynTree::FrameIndex>(10 + c), models.back());
    }
    auto system = std::make_shared<FloatingBaseDynamicalSystem>();
    auto rhoHandler = std::make_shared<StdImplementation>();
    rhoHandler->setParameter("rho", rho);
    if (!system->initalize(rhoHandler) || !system->setKinDyn(kinDyn)) return -2;
    if (regularization && !system->setMassMatrixRegularization(regularization, n, n)) return -3;

    for (std::size_t s = 0; s < n_systems; ++s)
    {
        for (std::size_t i = 0; i < n * n; ++i) kinDyn->mass.data()[i] = mass_matrices[s * n * n + i];
        for (int i = 0; i < 6; ++i) kinDyn->bias.baseWrench()(i) = bias_forces[s * n + i];
        for (std::size_t i = 0; i < dofs; ++i) kinDyn->bias.jointTorques()(i) = bias_forces[s * n + 6 + i];
        for (std::size_t c = 0; c < cps; ++c)
        {
            const std::size_t k = s * cps + c;
            const double* prm = params ? params + 4 * k : uniform;
            auto handler = std::make_shared<StdImplementation>();
            handler->setParameter("length", prm[0]);
            handler->setParameter("width", prm[1]);
            handler->setParameter("spring_coeff", prm[2]);
            handler->setParameter("damper_coeff", prm[3]);
            if (!models[c]->initialize(handler)) return -4;
            models[c]->setNullForceTransform(transformFromRow(null_poses + 12 * k));
            ArraysKinDyn::Frame& f = kinDyn->frames[contacts[c].index()];
            f.jacobian.resize(6, n);
            for (std::size_t i = 0; i < 6 * n; ++i) f.jacobian.data()[i] = jacobians[k * 6 * n + i];
            for (int i = 0; i < 6; ++i) f.velocity(i) = twists[6 * k + i];
            f.transform = transformFromRow(poses + 12 * k);
        }
        FloatingBaseDynamicalSystem::StateType x;
        auto& [v, sd, p, R, q] = x;
        sd = VectorXd(dofs);
        q = VectorXd(dofs);
        VectorXd tau(dofs);
        for (int i = 0; i < 6; ++i) v[i] = nu[s * n + i];
        for (std::size_t i = 0; i < dofs; ++i)
        {
            sd[i] = nu[s * n + 6 + i];
            q[i] = joint_pos ? joint_pos[s * dofs + i] : 0.0;
            tau[i] = joint_torques ? joint_torques[s * dofs + i] : 0.0;
        }
        for (int i = 0; i < 3; ++i) p[i] = base_pos[3 * s + i];
        fromRowMajor(base_rot + 9 * s, R);
        if (!system->setState(x) || !system->setControlInput({tau, contacts})) return -5;
        FloatingBaseDynamicalSystem::StateDerivativeType dx;
        if (!system->dynamics(0.0, dx)) return -6;
        for (int i = 0; i < 6; ++i) acc[s * n + i] = std::get<0>(dx)[i];
        for (std::size_t i = 0; i < dofs; ++i) acc[s * n + 6 + i] = std::get<1>(dx)[i];
        ForwardEuler<FloatingBaseDynamicalSystem> integrator(dT);
        if (!integrator.setDynamicalSystem(system) || !integrator.integrate(0.0, dT)) return -7;
        const auto& [v1, sd1, p1, R1, q1] = integrator.getSolution();
        for (int i = 0; i < 6; ++i) nu_out[s * n + i] = v1[i];
        for (std::size_t i = 0; i < dofs; ++i)
        {
            nu_out[s * n + 6 + i] = sd1[i];
            if (joint_pos_out) joint_pos_out[s * dofs + i] = q1[i];
        }
        for (int i = 0; i < 3; ++i) base_pos_out[3 * s + i] = p1[i];
        toRowMajor(R1, base_rot_out + 9 * s);
    }
    return 0;
}

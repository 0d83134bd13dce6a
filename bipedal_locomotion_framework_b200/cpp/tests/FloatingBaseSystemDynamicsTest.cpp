/**
 * @file FloatingBaseSystemDynamicsTest.cpp
 * System::FloatingBaseDynamicalSystem on the GPU (the reference has no test of this class; behaviour
 * from src/System/src/FloatingBaseSystemDynamics.cpp:17-251).  The KinDynComputations object is a test
 * double that answers with injected quantities -- the rigid-body algorithms are the caller's.
 *  - the protocol: every `return false` of the reference, in its order;
 *  - the numbers: M acc = -h + sum_c J_c^T wrench_c + [0; tau] with the wrenches of the per-instance
 *    contact models, with and without a regularisation term, with and without contacts, 23 and 0
 *    joints; linear velocity / rotation rate / joint velocity of the derivative;
 *  - ForwardEuler<FloatingBaseDynamicalSystem>: one integrate(0, dT) == x + dx dT.
 */
#ifdef BLF_HAVE_CATCH2
#include <catch2/catch.hpp>
#else
#include "catch_shim.h"
#endif

#include <chrono>
#include <cmath>
#include <cstdio>
#include <map>
#include <random>

#include <BipedalLocomotion/ContactModels/ContinuousContactModel.h>
#include <BipedalLocomotion/ContactModels/ContinuousContactModelBatch.h>
#include <BipedalLocomotion/ParametersHandler/StdImplementation.h>
#include <BipedalLocomotion/System/FloatingBaseSystemDynamics.h>
#include <BipedalLocomotion/System/FloatingBaseSystemKinematics.h>
#include <BipedalLocomotion/System/ForwardEuler.h>

using namespace BipedalLocomotion::ContactModels;
using namespace BipedalLocomotion::ParametersHandler;
using namespace BipedalLocomotion::System;

namespace
{
/** Answers with what the test injected; counts the state updates. */
class InjectedKinDyn : public iDynTree::KinDynComputations
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
    int stateUpdates{0};
    bool failState{false}, failMass{false}, failBias{false};
    iDynTree::Vector3 lastGravity;
    iDynTree::Transform lastBase;

    explicit InjectedKinDyn(std::size_t dofs) : robot(dofs), bias(robot) {}
    const iDynTree::Model& model() const override { return robot; }
    bool setRobotState(const iDynTree::Transform& base, const iDynTree::VectorDynSize& s, const iDynTree::Twist&,
                       const iDynTree::VectorDynSize& ds, const iDynTree::Vector3& gravity) override
    {
        ++stateUpdates;
        lastGravity = gravity;
        lastBase = base;
        return !failState && s.size() == robot.getNrOfDOFs() && ds.size() == robot.getNrOfDOFs();
    }
    bool getFreeFloatingMassMatrix(iDynTree::MatrixDynSize& out) override
    {
        if (failMass) return false;
        out = mass;
        return true;
    }
    bool generalizedBiasForces(iDynTree::FreeFloatingGeneralizedTorques& out) override
    {
        if (failBias) return false;
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

struct Problem
{
    std::shared_ptr<InjectedKinDyn> kinDyn;
    std::vector<std::shared_ptr<ContinuousContactModel>> models;
    std::vector<ContactWrench> contacts;
    FloatingBaseDynamicalSystem::StateType state;
    VectorXd torques;
};

Problem makeProblem(std::size_t dofs, int nContacts, unsigned seed)
{
    std::mt19937 gen(seed);
    std::uniform_real_distribution<> u(-1.0, 1.0);
    const std::size_t n = dofs + 6;
    Problem p;
    p.kinDyn = std::make_shared<InjectedKinDyn>(dofs);
    // symmetric positive definite mass matrix: A A^T / n + I / 2
    std::vector<double> A(n * n);
    for (double& x : A) x = u(gen);
    p.kinDyn->mass.resize(n, n);
    for (std::size_t i = 0; i < n; ++i)
        for (std::size_t k = 0; k <= i; ++k)
        {
            double v = (i == k) ? 0.5 : 0.0;
            for (std::size_t q = 0; q < n; ++q) v += A[i * n + q] * A[k * n + q] / static_cast<double>(n);
            p.kinDyn->mass(i, k) = p.kinDyn->mass(k, i) = v;
        }
    for (int i = 0; i < 6; ++i) p.kinDyn->bias.baseWrench()(i) = 20.0 * u(gen);
    for (std::size_t i = 0; i < dofs; ++i) p.kinDyn->bias.jointTorques()(i) = 20.0 * u(gen);
    for (int c = 0; c < nContacts; ++c)
    {
        InjectedKinDyn::Frame f;
        f.jacobian.resize(6, n);
        for (std::size_t i = 0; i < 6 * n; ++i) f.jacobian.data()[i] = u(gen);
        for (int i = 0; i < 6; ++i) f.velocity(i) = 0.3 * u(gen);
        f.transform = iDynTree::Transform(iDynTree::Rotation::RPY(0.2 * u(gen), 0.2 * u(gen), u(gen)),
                                          iDynTree::Position(u(gen), u(gen), 0.01 * u(gen)));
        const iDynTree::FrameIndex frame = 3 + 4 * c;
        p.kinDyn->frames[frame] = f;
        auto handler = std::make_shared<StdImplementation>();
        handler->setParameter("length", 0.12 + 0.01 * c);
        handler->setParameter("width", 0.09);
        handler->setParameter("spring_coeff", 2000.0 + 100.0 * c);
        handler->setParameter("damper_coeff", 100.0 - 10.0 * c);
        auto model = std::make_shared<ContinuousContactModel>();
        REQUIRE(model->initialize(handler));
        model->setNullForceTransform(iDynTree::Transform(iDynTree::Rotation::RPY(0, 0, u(gen)),
                                                         iDynTree::Position(u(gen), u(gen), 0.0)));
        p.models.push_back(model);
        p.contacts.emplace_back(frame, model);
    }
    auto& [baseVelocity, jointVelocity, basePosition, baseRotation, jointPositions] = p.state;
    for (int i = 0; i < 6; ++i) baseVelocity[i] = u(gen);
    jointVelocity = VectorXd(dofs);
    jointPositions = VectorXd(dofs);
    p.torques = VectorXd(dofs);
    for (std::size_t i = 0; i < dofs; ++i)
    {
        jointVelocity[i] = u(gen);
        jointPositions[i] = u(gen);
        p.torques[i] = 5.0 * u(gen);
    }
    for (int i = 0; i < 3; ++i) basePosition[i] = u(gen);
    const iDynTree::Rotation R = iDynTree::Rotation::RPY(u(gen), u(gen), u(gen));
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) baseRotation(r, c) = 1.01 * R(r, c);   // off the manifold: rho acts
    return p;
}

/** max_i |(M + reg) acc - rhs|_i / (|rhs terms| + |M| |acc|), rhs rebuilt from per-instance wrenches. */
double residual(const Problem& p, const std::vector<double>* reg, const FloatingBaseDynamicalSystem::StateDerivativeType& dx)
{
    const std::size_t dofs = p.kinDyn->robot.getNrOfDOFs(), n = dofs + 6;
    std::vector<double> acc(n), rhs(n), mag(n);
    for (int i = 0; i < 6; ++i) acc[i] = std::get<0>(dx)[i];
    for (std::size_t i = 0; i < dofs; ++i) acc[6 + i] = std::get<1>(dx)[i];
    for (int i = 0; i < 6; ++i) rhs[i] = -p.kinDyn->bias.baseWrench()(i);
    for (std::size_t i = 0; i < dofs; ++i) rhs[6 + i] = -p.kinDyn->bias.jointTorques()(i) + p.torques[i];
    for (std::size_t i = 0; i < n; ++i) mag[i] = std::fabs(rhs[i]) + (i >= 6 ? std::fabs(p.torques[i - 6]) : 0.0);
    for (std::size_t c = 0; c < p.contacts.size(); ++c)
    {
        const auto& f = p.kinDyn->frames.at(p.contacts[c].index());
        p.models[c]->setState(f.velocity, f.transform);
        const iDynTree::Wrench& w = p.models[c]->getContactWrench();
        for (std::size_t q = 0; q < n; ++q)
            for (int r = 0; r < 6; ++r)
            {
                rhs[q] += f.jacobian(r, q) * w(r);
                mag[q] += std::fabs(f.jacobian(r, q) * w(r));
            }
    }
    double m = 0, xm = 0, worst = 0;
    for (std::size_t i = 0; i < n; ++i) m = std::max(m, mag[i]);
    for (std::size_t i = 0; i < n; ++i) xm = std::max(xm, std::fabs(acc[i]));
    for (std::size_t i = 0; i < n; ++i)
    {
        double res = -rhs[i], rowSum = 0;
        for (std::size_t k = 0; k < n; ++k)
        {
            const double a = p.kinDyn->mass(i, k) + (reg ? (*reg)[i * n + k] : 0.0);
            res += a * acc[k];
            rowSum += std::fabs(a);
        }
        worst = std::max(worst, std::fabs(res) / (m + rowSum * xm));
    }
    return worst;
}
} // namespace

TEST_CASE("FloatingBaseDynamicalSystem protocol")
{
    FloatingBaseDynamicalSystem system;
    FloatingBaseDynamicalSystem::StateDerivativeType dx;
    REQUIRE_FALSE(system.dynamics(0.0, dx));                                   // no KinDyn yet
    REQUIRE_FALSE(system.initalize(std::weak_ptr<IParametersHandler>()));      // expired handler
    auto handler = std::make_shared<StdImplementation>();
    REQUIRE_FALSE(system.initalize(handler));                                  // no "rho"
    handler->setParameter("rho", 0.5);
    REQUIRE(system.initalize(handler));
    REQUIRE_FALSE(system.setKinDyn(nullptr));
    std::vector<double> reg(29 * 29, 0.0);
    REQUIRE_FALSE(system.setMassMatrixRegularization(reg.data(), 29, 29));     // before setKinDyn

    Problem p = makeProblem(23, 2, 11);
    REQUIRE(system.setKinDyn(p.kinDyn));
    REQUIRE_FALSE(system.setMassMatrixRegularization(reg.data(), 28, 28));     // wrong size
    REQUIRE_FALSE(system.setMassMatrixRegularization(reg.data(), 29, 28));
    REQUIRE(system.setState(p.state));
    REQUIRE(system.setControlInput({VectorXd(22), p.contacts}));
    REQUIRE_FALSE(system.dynamics(0.0, dx));                                   // wrong size of the torques
    REQUIRE(p.kinDyn->stateUpdates == 0);
    REQUIRE(system.setControlInput({p.torques, p.contacts}));
    p.kinDyn->failState = true;
    REQUIRE_FALSE(system.dynamics(0.0, dx));
    p.kinDyn->failState = false;
    p.kinDyn->failMass = true;
    REQUIRE_FALSE(system.dynamics(0.0, dx));
    p.kinDyn->failMass = false;
    p.kinDyn->failBias = true;
    REQUIRE_FALSE(system.dynamics(0.0, dx));
    p.kinDyn->failBias = false;
    {   // a frame the KinDyn object does not know; a contact model that has expired
        std::vector<ContactWrench> unknown{ContactWrench(iDynTree::FrameIndex(99), p.models[0])};
        REQUIRE(system.setControlInput({p.torques, unknown}));
        REQUIRE_FALSE(system.dynamics(0.0, dx));
        std::vector<ContactWrench> expired{ContactWrench(p.contacts[0].index(), nullptr)};
        REQUIRE(system.setControlInput({p.torques, expired}));
        REQUIRE_FALSE(system.dynamics(0.0, dx));
    }
    REQUIRE(system.setControlInput({p.torques, p.contacts}));
    const int before = p.kinDyn->stateUpdates;
    REQUIRE(system.dynamics(0.0, dx));
    REQUIRE(p.kinDyn->stateUpdates == before + 1);
    REQUIRE(p.kinDyn->lastGravity(2) == -9.81);                                // the constructor's gravity
    system.setGravityVector(Vector3d{0.0, 0.0, -1.62});
    REQUIRE(system.dynamics(0.0, dx));
    REQUIRE(p.kinDyn->lastGravity(2) == -1.62);
    for (int k = 0; k < 3; ++k) REQUIRE(p.kinDyn->lastBase.getPosition()(k) == std::get<2>(p.state)[k]);
}

TEST_CASE("FloatingBaseDynamicalSystem dynamics")
{
    SECTION("23 joints, two contacts, with and without a regularisation term")
    {
        Problem p = makeProblem(23, 2, 5);
        auto handler = std::make_shared<StdImplementation>();
        handler->setParameter("rho", 0.7);
        FloatingBaseDynamicalSystem system;
        REQUIRE(system.initalize(handler));
        REQUIRE(system.setKinDyn(p.kinDyn));
        REQUIRE(system.setState(p.state));
        REQUIRE(system.setControlInput({p.torques, p.contacts}));
        FloatingBaseDynamicalSystem::StateDerivativeType dx;
        REQUIRE(system.dynamics(0.0, dx));
        REQUIRE(residual(p, nullptr, dx) <= 1e-12);
        REQUIRE(std::get<1>(dx).size() == 23);

        // the rest of the derivative: linear velocity, joint velocity, and the rotation rate of the
        // kinematics' dynamics() for the same rotation and twist (same device formula: equal)
        for (int k = 0; k < 3; ++k) REQUIRE(std::get<2>(dx)[k] == std::get<0>(p.state)[k]);
        for (int k = 0; k < 23; ++k) REQUIRE(std::get<4>(dx)[k] == std::get<1>(p.state)[k]);
        auto kinematics = std::make_shared<FloatingBaseSystemKinematics>();
        REQUIRE(kinematics->initalize(handler));
        REQUIRE(kinematics->setState({std::get<2>(p.state), std::get<3>(p.state), std::get<4>(p.state)}));
        REQUIRE(kinematics->setControlInput({std::get<0>(p.state), std::get<1>(p.state)}));
        FloatingBaseSystemKinematics::StateDerivativeType kdx;
        REQUIRE(kinematics->dynamics(0.0, kdx));
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) REQUIRE(std::get<3>(dx)(r, c) == std::get<1>(kdx)(r, c));

        // the contact models were handed the frame states (the reference's side effect)
        const auto& f = p.kinDyn->frames.at(p.contacts[1].index());
        double st[30], prm[4];
        p.models[1]->batchInputs(st, prm);
        REQUIRE(prm[0] == 0.13);
        REQUIRE(prm[3] == 90.0);
        for (int k = 0; k < 6; ++k) REQUIRE(st[k] == f.velocity(k));
        for (int k = 0; k < 3; ++k) REQUIRE(st[6 + k] == f.transform.getPosition()(k));

        {   // what a call costs (one upload, two launches, one download + the rotation-rate call)
            const auto t0 = std::chrono::steady_clock::now();
            const int calls = 200;
            for (int i = 0; i < calls; ++i) REQUIRE(system.dynamics(0.0, dx));
            const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
            std::printf("FloatingBaseDynamicalSystem::dynamics, 23 joints, 2 contacts: %.1f us per call\n", us / calls);
        }

        // regularisation: (M + reg) acc = rhs
        const std::size_t n = 29;
        std::vector<double> reg(n * n, 0.0);
        for (std::size_t i = 0; i < n; ++i) reg[i * n + i] = 0.1 + 0.01 * static_cast<double>(i);
        REQUIRE(system.setMassMatrixRegularization(reg.data(), n, n));
        FloatingBaseDynamicalSystem::StateDerivativeType dxReg;
        REQUIRE(system.dynamics(0.0, dxReg));
        REQUIRE(residual(p, &reg, dxReg) <= 1e-12);
        REQUIRE(std::fabs(std::get<0>(dxReg)[0] - std::get<0>(dx)[0]) > 1e-6);   // it did change the answer
    }

    SECTION("no contacts; no joints")
    {
        Problem p = makeProblem(23, 0, 6);
        FloatingBaseDynamicalSystem system;
        REQUIRE(system.setKinDyn(p.kinDyn));
        REQUIRE(system.setState(p.state));
        REQUIRE(system.setControlInput({p.torques, p.contacts}));
        FloatingBaseDynamicalSystem::StateDerivativeType dx;
        REQUIRE(system.dynamics(0.0, dx));
        REQUIRE(residual(p, nullptr, dx) <= 1e-12);

        Problem q = makeProblem(0, 1, 8);
        FloatingBaseDynamicalSystem rigidBody;
        REQUIRE(rigidBody.setKinDyn(q.kinDyn));
        REQUIRE(rigidBody.setState(q.state));
        REQUIRE(rigidBody.setControlInput({q.torques, q.contacts}));
        REQUIRE(rigidBody.dynamics(0.0, dx));
        REQUIRE(std::get<1>(dx).size() == 0);
        REQUIRE(residual(q, nullptr, dx) <= 1e-12);
    }

    SECTION("ForwardEuler: one step is x + dx dT")
    {
        Problem p = makeProblem(12, 2, 9);
        auto system = std::make_shared<FloatingBaseDynamicalSystem>();
        REQUIRE(system->setKinDyn(p.kinDyn));
        REQUIRE(system->setState(p.state));
        REQUIRE(system->setControlInput({p.torques, p.contacts}));
        FloatingBaseDynamicalSystem::StateDerivativeType dx;
        REQUIRE(system->dynamics(0.0, dx));
        const double dT = 0.01;
        ForwardEuler<FloatingBaseDynamicalSystem> integrator(dT);
        REQUIRE(integrator.setDynamicalSystem(system));
        REQUIRE(integrator.integrate(0, dT));
        const auto& [v, sd, pos, R, s] = integrator.getSolution();
        for (int k = 0; k < 6; ++k) REQUIRE(v[k] == std::get<0>(p.state)[k] + std::get<0>(dx)[k] * dT);
        for (int k = 0; k < 12; ++k) REQUIRE(sd[k] == std::get<1>(p.state)[k] + std::get<1>(dx)[k] * dT);
        for (int k = 0; k < 3; ++k) REQUIRE(pos[k] == std::get<2>(p.state)[k] + std::get<2>(dx)[k] * dT);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                REQUIRE(R(r, c) == std::get<3>(p.state)(r, c) + std::get<3>(dx)(r, c) * dT);
        for (int k = 0; k < 12; ++k) REQUIRE(s[k] == std::get<4>(p.state)[k] + std::get<4>(dx)[k] * dT);
    }
}

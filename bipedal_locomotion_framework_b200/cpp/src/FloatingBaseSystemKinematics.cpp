/**
 * @file FloatingBaseSystemKinematics.cpp
 * Facade over blf_sys_kinematics_dynamics_host / blf_sys_kinematics_integrate_host
 * (reference: src/System/src/FloatingBaseSystemKinematics.cpp:13-73).
 */
#include <iostream>

#include <BipedalLocomotion/ContactModels/ContinuousContactModelBatch.h>
#include <BipedalLocomotion/System/FloatingBaseSystemKinematics.h>

#include "blf_ccm.h"

using namespace BipedalLocomotion::System;
using namespace BipedalLocomotion::ParametersHandler;
using BipedalLocomotion::ContactModels::CudaDevice;

bool FloatingBaseSystemKinematics::ensureDevice(const char* where)
{
    if (m_device == nullptr) m_device = CudaDevice::open(m_deviceIndex);
    if (m_device == nullptr)
    {
        std::cerr << "[FloatingBaseSystemKinematics::" << where
                  << "] The CUDA backend is not available and there is no CPU evaluation path." << std::endl;
        return false;
    }
    return true;
}

bool FloatingBaseSystemKinematics::initalize(std::weak_ptr<IParametersHandler> handler)
{
    auto ptr = handler.lock();
    if (ptr == nullptr)
    {
        std::cerr << "[FloatingBaseSystemKinematics::initalize] The parameter handler is expired. "
                     "Please call the function passing a pointer pointing an already allocated "
                     "memory."
                  << std::endl;
        return false;
    }
    if (!ptr->getParameter("rho", m_rho))
    {
        std::cerr << "[FloatingBaseSystemKinematics::initalize] Unable to load the Baumgarte "
                     "stabilization parameter."
                  << std::endl;
        return false;
    }
    return true;
}

bool FloatingBaseSystemKinematics::dynamics(const double& /*time*/, StateDerivativeType& stateDerivative)
{
    const Matrix3d& baseRotation = std::get<1>(m_state);
    const VectorXd& jointPositions = std::get<2>(m_state);
    const Vector6d& baseTwist = std::get<0>(m_controlInput);
    const VectorXd& jointVelocity = std::get<1>(m_controlInput);

    if (jointVelocity.size() != jointPositions.size())
    {
        std::cerr << "[FloatingBaseSystemKinematics::dynamics] Wrong size of the vectors." << std::endl;
        return false;
    }
    if (!ensureDevice("dynamics")) return false;

    Vector3d& baseLinearVelocity = std::get<0>(stateDerivative);
    double rotation[9], rotationRate[9]; // row-major across the C ABI
    toRowMajor(baseRotation, rotation);
    const int rc = blf_sys_kinematics_dynamics_host(static_cast<blf_ccm_handle*>(m_device->handle()), 1,
                                                    m_rho, baseTwist.data(), rotation,
                                                    baseLinearVelocity.data(), rotationRate);
    if (rc != BLF_CCM_OK)
    {
        std::cerr << "[FloatingBaseSystemKinematics::dynamics] " << blf_ccm_last_error() << std::endl;
        return false;
    }
    fromRowMajor(rotationRate, std::get<1>(stateDerivative));
    std::get<2>(stateDerivative) = jointVelocity;
    return true;
}

bool FloatingBaseSystemKinematics::advanceOnDevice(double stepDT, double lastDT, int steps)
{
    Vector3d& basePosition = std::get<0>(m_state);
    Matrix3d& baseRotation = std::get<1>(m_state);
    VectorXd& jointPositions = std::get<2>(m_state);
    const Vector6d& baseTwist = std::get<0>(m_controlInput);
    const VectorXd& jointVelocity = std::get<1>(m_controlInput);

    if (jointVelocity.size() != jointPositions.size())
    {
        std::cerr << "[FloatingBaseSystemKinematics::dynamics] Wrong size of the vectors." << std::endl;
        return false;
    }
    if (!ensureDevice("dynamics")) return false;
    const std::int64_t nj = static_cast<std::int64_t>(jointPositions.size());
    double rotation[9]; // row-major across the C ABI
    toRowMajor(baseRotation, rotation);
    const int rc = blf_sys_kinematics_integrate_host(static_cast<blf_ccm_handle*>(m_device->handle()), 1,
                                                     m_rho, stepDT, lastDT, steps, baseTwist.data(),
                                                     basePosition.data(), rotation, nj,
                                                     nj ? jointVelocity.data() : nullptr,
                                                     nj ? jointPositions.data() : nullptr);
    if (rc != BLF_CCM_OK)
    {
        std::cerr << "[FloatingBaseSystemKinematics::dynamics] " << blf_ccm_last_error() << std::endl;
        return false;
    }
    fromRowMajor(rotation, baseRotation);
    return true;
}

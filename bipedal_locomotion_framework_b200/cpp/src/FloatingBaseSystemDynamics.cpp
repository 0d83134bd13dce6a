/**
 * @file FloatingBaseSystemDynamics.cpp
 * Facade over blf_sys_floating_base_acceleration / blf_sys_mass_matrix_solve /
 * blf_sys_kinematics_dynamics with one system (reference:
 * src/System/src/FloatingBaseSystemDynamics.cpp:17-251).  The order of the checks, their messages and
 * the calls made on the KinDynComputations object and on the contact models follow the reference; the
 * arithmetic is the device's.
 */
#include <cstring>
#include <iostream>

#include <BipedalLocomotion/ContactModels/ContinuousContactModel.h>
#include <BipedalLocomotion/ContactModels/ContinuousContactModelBatch.h>
#include <BipedalLocomotion/System/FloatingBaseSystemDynamics.h>

#include "blf_ccm.h"

using namespace BipedalLocomotion::System;
using namespace BipedalLocomotion::ParametersHandler;
using BipedalLocomotion::ContactModels::ContinuousContactModel;
using BipedalLocomotion::ContactModels::CudaDevice;
using BipedalLocomotion::GenericContainer::DeviceSoA;

FloatingBaseDynamicalSystem::FloatingBaseDynamicalSystem()
{
    m_gravity.zero();
    m_gravity(2) = -9.81;
}

FloatingBaseDynamicalSystem::FloatingBaseDynamicalSystem(int device) : FloatingBaseDynamicalSystem()
{
    m_deviceIndex = device;
}

FloatingBaseDynamicalSystem::FloatingBaseDynamicalSystem(std::shared_ptr<CudaDevice> device)
    : FloatingBaseDynamicalSystem()
{
    m_device = std::move(device);
}

bool FloatingBaseDynamicalSystem::ensureDevice(const char* where)
{
    if (m_device == nullptr)
        m_device = CudaDevice::open(m_deviceIndex >= 0 ? m_deviceIndex : CudaDevice::defaultIndex());
    if (m_device == nullptr)
    {
        std::cerr << "[FloatingBaseDynamicalSystem::" << where
                  << "] The CUDA backend is not available and there is no CPU evaluation path." << std::endl;
        return false;
    }
    return true;
}

bool FloatingBaseDynamicalSystem::initalize(std::weak_ptr<IParametersHandler> handler)
{
    auto ptr = handler.lock();
    if (ptr == nullptr)
    {
        std::cerr << "[FloatingBaseDynamicalSystem::initalize] The parameter handler is expired. "
                     "Please call the function passing a pointer pointing an already allocated "
                     "memory."
                  << std::endl;
        return false;
    }
    if (!ptr->getParameter("rho", m_rho))
    {
        std::cerr << "[FloatingBaseDynamicalSystem::initalize] Unable to load the Baumgarte "
                     "stabilization parameter."
                  << std::endl;
        return false;
    }
    return true;
}

void FloatingBaseDynamicalSystem::setGravityVector(const Vector3d& gravity)
{
    for (int i = 0; i < 3; ++i) m_gravity(i) = gravity[i];
}

bool FloatingBaseDynamicalSystem::setKinDyn(std::shared_ptr<iDynTree::KinDynComputations> kinDyn)
{
    if (kinDyn == nullptr)
    {
        std::cerr << "[FloatingBaseDynamicalSystem::setKinDynComputation] Corrupted KinDyn "
                     "computation object."
                  << std::endl;
        return false;
    }
    m_kinDyn = kinDyn;
    m_actuatedDoFs = m_kinDyn->model().getNrOfDOFs();
    m_massMatrix.resize(m_actuatedDoFs + m_baseDoFs, m_actuatedDoFs + m_baseDoFs);
    m_jacobianMatrix.resize(m_baseDoFs, m_actuatedDoFs + m_baseDoFs);
    m_generalizedBiasForces.resize(m_kinDyn->model());
    m_acceleration.assign(m_actuatedDoFs + m_baseDoFs, 0.0);
    return true;
}

bool FloatingBaseDynamicalSystem::setMassMatrixRegularization(const double* matrix, std::size_t rows,
                                                              std::size_t cols)
{
    if (m_kinDyn == nullptr)
    {
        std::cerr << "[FloatingBaseDynamicalSystem::setMassMatrixRegularization] Please call "
                     "'setKinDyn()' before."
                  << std::endl;
        return false;
    }
    if ((m_actuatedDoFs + m_baseDoFs != rows) || (cols != rows) || matrix == nullptr)
    {
        const auto rightSize = m_actuatedDoFs + m_baseDoFs;
        std::cerr << "[FloatingBaseDynamicalSystem::setMassMatrixRegularization] The size of the "
                     "regularization matrix is not correct. The correct size is: "
                  << rightSize << " x " << rightSize << ". While the input of the function is a " << rows
                  << " x " << cols << " matrix." << std::endl;
        return false;
    }
    m_massMatrixReglarizationTerm.assign(matrix, matrix + rows * cols);
    m_useMassMatrixRegularizationTerm = true;
    return true;
}

bool FloatingBaseDynamicalSystem::setMassMatrixRegularization(const iDynTree::MatrixDynSize& matrix)
{
    return setMassMatrixRegularization(matrix.data(), matrix.rows(), matrix.cols());
}

bool FloatingBaseDynamicalSystem::dynamics(const double& /*time*/, StateDerivativeType& stateDerivative)
{
    if (m_kinDyn == nullptr)
    {
        std::cerr << "[FloatingBaseDynamicalSystem::dynamics] Please call 'setKinDyn()' before." << std::endl;
        return false;
    }

    const auto& [baseVelocity, jointVelocity, basePosition, baseOrientation, jointPositions] = m_state;
    auto& [baseAcceleration, jointAcceleration, baseLinearVelocity, baseRotationRate, jointVelocityOutput]
        = stateDerivative;
    const VectorXd& jointTorques = std::get<0>(m_controlInput);
    const std::vector<ContactWrench>& contactWrenches = std::get<1>(m_controlInput);

    if (static_cast<std::size_t>(jointVelocity.size()) != m_actuatedDoFs
        || static_cast<std::size_t>(jointPositions.size()) != m_actuatedDoFs
        || static_cast<std::size_t>(jointTorques.size()) != m_actuatedDoFs)
    {
        std::cerr << "[FloatingBaseDynamicalSystem::dynamics] Wrong size of the vectors." << std::endl;
        return false;
    }
    if (!ensureDevice("dynamics")) return false;
    blf_ccm_handle* h = static_cast<blf_ccm_handle*>(m_device->handle());

    // base twist and rotation as they cross the C ABI; linear velocity and rotation rate (:134-140) come
    // back from the device with the acceleration
    double twist[6], rotation[9];
    for (int i = 0; i < 6; ++i) twist[i] = baseVelocity[i];
    toRowMajor(baseOrientation, rotation);
    jointVelocityOutput = jointVelocity;

    // update the kinDynComputations object (:144-170)
    iDynTree::Twist baseTwist;
    for (int i = 0; i < 6; ++i) baseTwist(i) = baseVelocity[i];
    iDynTree::Rotation baseRot;
    std::memcpy(baseRot.data(), rotation, sizeof(rotation));
    iDynTree::Position basePos(basePosition[0], basePosition[1], basePosition[2]);
    iDynTree::VectorDynSize jointPos(m_actuatedDoFs), jointVel(m_actuatedDoFs);
    for (std::size_t i = 0; i < m_actuatedDoFs; ++i)
    {
        jointPos(i) = jointPositions[i];
        jointVel(i) = jointVelocity[i];
    }
    if (!m_kinDyn->setRobotState(iDynTree::Transform(baseRot, basePos), jointPos, baseTwist, jointVel, m_gravity))
    {
        std::cerr << "[FloatingBaseDynamicalSystem::dynamics] Unable to update the kindyn object." << std::endl;
        return false;
    }
    if (!m_kinDyn->getFreeFloatingMassMatrix(m_massMatrix))
    {
        std::cerr << "[FloatingBaseDynamicalSystem::dynamics] Unable to get the mass matrix." << std::endl;
        return false;
    }
    if (!m_kinDyn->generalizedBiasForces(m_generalizedBiasForces))
    {
        std::cerr << "[FloatingBaseDynamicalSystem::dynamics] Unable to get the bias forces." << std::endl;
        return false;
    }

    // the device block: [bias | torques | mass | regularization | states 30 x C | parameters 4 x C |
    // Jacobians C x 6 x n | acceleration], every piece an even number of doubles from the start
    const std::size_t n = m_actuatedDoFs + m_baseDoFs, contacts = contactWrenches.size();
    if (n > 128)
    {
        std::cerr << "[FloatingBaseDynamicalSystem::dynamics] " << n
                  << " unknowns: the CUDA backend solves up to 128." << std::endl;
        return false;
    }
    auto even = [](std::size_t x) { return (x + 1) & ~static_cast<std::size_t>(1); };
    const std::size_t atBias = 0, atTau = atBias + even(n), atMass = atTau + even(m_actuatedDoFs),
                      atReg = atMass + even(n * n),
                      atStates = atReg + (m_useMassMatrixRegularizationTerm ? even(n * n) : 0),
                      atParams = atStates + even(30 * contacts), atJac = atParams + even(4 * contacts),
                      atTwist = atJac + even(contacts * 6 * n), atRot = atTwist + 6,
                      atAcc = atRot + 10,   // results, one run: acceleration | linear velocity (3) | rotation rate (9)
                      atLin = atAcc + even(n), atRate = atLin + 4, total = atRate + 10;
    m_staging.assign(total, 0.0);
    double* s = m_staging.data();
    std::memcpy(s + atBias, m_generalizedBiasForces.baseWrench().data(), 6 * sizeof(double));
    for (std::size_t i = 0; i < m_actuatedDoFs; ++i)
    {
        s[atBias + 6 + i] = m_generalizedBiasForces.jointTorques()(i);
        s[atTau + i] = jointTorques[i];
    }
    std::memcpy(s + atMass, m_massMatrix.data(), n * n * sizeof(double));
    std::memcpy(s + atTwist, twist, sizeof(twist));
    std::memcpy(s + atRot, rotation, sizeof(rotation));
    if (m_useMassMatrixRegularizationTerm)
        std::memcpy(s + atReg, m_massMatrixReglarizationTerm.data(), n * n * sizeof(double));

    // the contacts (:199-226): Jacobian, state of the contact model; their wrenches are evaluated by
    // the device call below, all at once
    for (std::size_t c = 0; c < contacts; ++c)
    {
        const ContactWrench& contactWrench = contactWrenches[c];
        if (!m_kinDyn->getFrameFreeFloatingJacobian(contactWrench.index(), m_jacobianMatrix))
        {
            std::cerr << "[FloatingBaseDynamicalSystem::dynamics] Unable to get the Jacobian for "
                         "the frame named: "
                      << m_kinDyn->model().getFrameLink(contactWrench.index()) << "." << std::endl;
            return false;
        }
        auto contactPtr = contactWrench.contactModel().lock();
        if (contactPtr == nullptr)
        {
            std::cerr << "[FloatingBaseDynamicalSystem::dynamics] The contact model associated to "
                         "the frame named: "
                      << m_kinDyn->model().getFrameLink(contactWrench.index()) << " has been expired."
                      << std::endl;
            return false;
        }
        contactPtr->setState(m_kinDyn->getFrameVel(contactWrench.index()),
                             m_kinDyn->getWorldTransform(contactWrench.index()));
        const auto* continuous = dynamic_cast<const ContinuousContactModel*>(contactPtr.get());
        if (continuous == nullptr)
        {
            std::cerr << "[FloatingBaseDynamicalSystem::dynamics] The contact model associated to "
                         "the frame named: "
                      << m_kinDyn->model().getFrameLink(contactWrench.index())
                      << " is not a ContinuousContactModel: the CUDA backend evaluates no other model."
                      << std::endl;
            return false;
        }
        if (m_jacobianMatrix.rows() != m_baseDoFs || m_jacobianMatrix.cols() != n)
        {
            std::cerr << "[FloatingBaseDynamicalSystem::dynamics] Wrong size of the Jacobian." << std::endl;
            return false;
        }
        double state[30], parameters[4];
        continuous->batchInputs(state, parameters);
        for (int k = 0; k < 30; ++k) s[atStates + k * contacts + c] = state[k];       // planes
        for (int k = 0; k < 4; ++k) s[atParams + k * contacts + c] = parameters[k];
        std::memcpy(s + atJac + c * 6 * n, m_jacobianMatrix.data(), 6 * n * sizeof(double));
    }
    if (contacts == 0)   // known = -h: the sign change is exact
        for (std::size_t i = 0; i < n; ++i) s[atBias + i] = -s[atBias + i];

    // one upload (everything up to the results), the launches, one download, one synchronisation
    if (!m_deviceBlock.valid() || m_deviceBlock.size() != total) m_deviceBlock = DeviceSoA(m_device, 1, total);
    if (!m_deviceBlock.valid()
        || blf_ccm_copy_h2d(h, m_deviceBlock.plane(0), s, atAcc * sizeof(double), nullptr) != BLF_CCM_OK)
    {
        std::cerr << "[FloatingBaseDynamicalSystem::dynamics] " << blf_ccm_last_error() << std::endl;
        return false;
    }
    double* d = m_deviceBlock.plane(0);
    const double* tau = m_actuatedDoFs > 0 ? d + atTau : nullptr;
    const double* reg = m_useMassMatrixRegularizationTerm ? d + atReg : nullptr;
    int rc;
    if (contacts > 0)
    {
        const double* statePlanes[30];
        const double* parameterPlanes[4];
        for (int k = 0; k < 30; ++k) statePlanes[k] = d + atStates + k * contacts;
        for (int k = 0; k < 4; ++k) parameterPlanes[k] = d + atParams + k * contacts;
        rc = blf_sys_floating_base_acceleration(h, 1, static_cast<int>(contacts), static_cast<int>(n), statePlanes,
                                                parameterPlanes, d + atJac, d + atBias, tau, d + atMass, reg,
                                                d + atAcc, nullptr, nullptr);
    } else
    {
        rc = blf_sys_mass_matrix_solve(h, 1, static_cast<int>(n), d + atMass, reg, d + atBias, tau, d + atAcc, nullptr);
    }
    if (rc == BLF_CCM_OK)
        rc = blf_sys_kinematics_dynamics(h, 1, m_rho, d + atTwist, d + atRot, d + atLin, d + atRate, nullptr);
    m_acceleration.resize(total - atAcc);
    if (rc == BLF_CCM_OK)
        rc = blf_ccm_copy_d2h(h, m_acceleration.data(), d + atAcc, (total - atAcc) * sizeof(double), nullptr);
    if (rc == BLF_CCM_OK) rc = blf_ccm_stream_synchronize(h, nullptr);
    if (rc != BLF_CCM_OK)
    {
        std::cerr << "[FloatingBaseDynamicalSystem::dynamics] " << blf_ccm_last_error() << std::endl;
        return false;
    }

    for (int i = 0; i < 3; ++i) baseLinearVelocity[i] = m_acceleration[atLin - atAcc + i];
    fromRowMajor(m_acceleration.data() + (atRate - atAcc), baseRotationRate);

    // split the acceleration in base and joint acceleration (:245-248)
    for (std::size_t i = 0; i < m_baseDoFs; ++i) baseAcceleration[i] = m_acceleration[i];
    jointAcceleration.resize(m_actuatedDoFs);
    for (std::size_t i = 0; i < m_actuatedDoFs; ++i) jointAcceleration[i] = m_acceleration[m_baseDoFs + i];
    return true;
}

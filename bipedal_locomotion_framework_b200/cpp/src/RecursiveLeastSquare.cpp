/**
 * @file RecursiveLeastSquare.cpp
 * Host side of the GPU recursive-least-squares estimator; the arithmetic is in
 * csrc/rls_kernels.cuh.  Parameter handling and messages follow the reference
 * (src/Estimators/src/RecursiveLeastSquare.cpp:17-149).
 */
#include <cassert>
#include <iostream>

#include <BipedalLocomotion/Estimators/RecursiveLeastSquare.h>

#include "blf_ccm.h"

using namespace BipedalLocomotion::Estimators;
using namespace BipedalLocomotion::ParametersHandler;
using BipedalLocomotion::ContactModels::CudaDevice;
using BipedalLocomotion::GenericContainer::DeviceSoA;

namespace
{
blf_ccm_handle* raw(const std::shared_ptr<CudaDevice>& d)
{
    return d ? static_cast<blf_ccm_handle*>(d->handle()) : nullptr;
}
} // namespace

bool RecursiveLeastSquare::initialize(std::weak_ptr<IParametersHandler> handlerWeak)
{
    if (m_estimatorState != State::NotInitialized)
    {
        std::cerr << "[RecursiveLeastSquare::initialize] The estimator has been already initialized."
                  << std::endl;
        return false;
    }
    auto handler = handlerWeak.lock();
    if (handler == nullptr)
    {
        std::cerr << "[RecursiveLeastSquare::initialize] The parameter handler is expired. Please "
                     "check its scope."
                  << std::endl;
        return false;
    }
    if (!handler->getParameter("measurement_covariance", m_measurementCovariance, BipedalLocomotion::GenericContainer::VectorResizeMode::Resizable))
    {
        std::cerr << "[RecursiveLeastSquare::initialize] Unable to find the covariance matrix of "
                     "the measuraments."
                  << std::endl;
        return false;
    }
    if (!handler->getParameter("lambda", m_lambda))
    {
        std::cerr << "[RecursiveLeastSquare::initialize] Unable to find lambda." << std::endl;
        return false;
    }
    std::vector<double> state, stateCovariance;
    if (!handler->getParameter("state", state, BipedalLocomotion::GenericContainer::VectorResizeMode::Resizable))
    {
        std::cerr << "[RecursiveLeastSquare::initialize] Unable to get the initial guess." << std::endl;
        return false;
    }
    if (!handler->getParameter("state_covariance", stateCovariance, BipedalLocomotion::GenericContainer::VectorResizeMode::Resizable))
    {
        std::cerr << "[RecursiveLeastSquare::initialize] Unable to get the initial state covariance."
                  << std::endl;
        return false;
    }
    if (stateCovariance.size() != state.size())
    {
        std::cerr << "[RecursiveLeastSquare::initialize] state and state_covariance differ in size."
                  << std::endl;
        return false;
    }
    m_device = CudaDevice::open(CudaDevice::defaultIndex());
    if (m_device == nullptr)
    {
        std::cerr << "[RecursiveLeastSquare::initialize] The CUDA backend is not available and there "
                     "is no CPU evaluation path."
                  << std::endl;
        return false;
    }
    m_state.resize(state.size());
    m_stateCovarianceMatrix.resize(state.size(), state.size());
    for (std::size_t i = 0; i < state.size(); ++i)
    {
        m_state(i) = state[i];
        m_stateCovarianceMatrix(i, i) = stateCovariance[i];
    }
    m_measurements.resize(m_measurementCovariance.size());
    m_measurements.zero();
    m_estimatorState = State::Initialized;
    return true;
}

void RecursiveLeastSquare::setRegressorFunction(std::function<iDynTree::MatrixDynSize(void)> regressor)
{
    m_regressor = regressor;
}

bool RecursiveLeastSquare::advance()
{
    if (m_regressor == nullptr)
    {
        std::cerr << "[RecursiveLeastSquare::advance] Please call the setRegressorFunction() before "
                     "calling advance"
                  << std::endl;
        return false;
    }
    if (m_estimatorState != State::Initialized && m_estimatorState != State::Running)
    {
        std::cerr << "[RecursiveLeastSquare::advance] Please initialize the estimator before calling "
                     "advance."
                  << std::endl;
        return false;
    }
    if (m_estimatorState == State::Initialized) m_estimatorState = State::Running;

    const iDynTree::MatrixDynSize regressor = m_regressor();
    if (regressor.rows() != m_measurements.size() || regressor.cols() != m_state.size())
    {
        std::cerr << "[RecursiveLeastSquare::advance] The regressor must be measurements x parameters."
                  << std::endl;
        return false;
    }
    const int rc = blf_rls_advance_host(raw(m_device), 1, static_cast<int>(m_state.size()),
                                        static_cast<int>(m_measurements.size()), regressor.data(),
                                        m_measurements.data(), m_measurementCovariance.data(), m_lambda,
                                        m_state.data(), m_stateCovarianceMatrix.data());
    if (rc != BLF_CCM_OK)
    {
        std::cerr << "[RecursiveLeastSquare::advance] " << blf_ccm_last_error() << std::endl;
        return false;
    }
    return true;
}

void RecursiveLeastSquare::setMeasurements(const iDynTree::VectorDynSize& measurements)
{
    assert(m_measurements.size() == measurements.size());
    m_measurements = measurements;
}

const iDynTree::VectorDynSize& RecursiveLeastSquare::parametersExpectedValue() const { return m_state; }

const iDynTree::MatrixDynSize& RecursiveLeastSquare::parametersCovarianceMatrix() const
{
    return m_stateCovarianceMatrix;
}

RecursiveLeastSquareBatch::RecursiveLeastSquareBatch(std::shared_ptr<CudaDevice> device,
                                                     std::vector<double> measurementCovariance,
                                                     double lambda)
    : m_device(std::move(device)), m_measurementCovariance(std::move(measurementCovariance)),
      m_lambda(lambda)
{
}

bool RecursiveLeastSquareBatch::advance(const DeviceSoA& regressor, const DeviceSoA& measurements,
                                        DeviceSoA& state, DeviceSoA& covariance, void* stream)
{
    const std::size_t p = state.planes(), m = measurements.planes();
    if (m_device == nullptr || m != m_measurementCovariance.size() || regressor.planes() != m * p
        || covariance.planes() != p * p)
    {
        std::cerr << "[RecursiveLeastSquareBatch::advance] Inconsistent plane counts." << std::endl;
        return false;
    }
    const int rc = blf_rls_advance_batch(raw(m_device), static_cast<std::int64_t>(state.size()),
                                         static_cast<int>(p), static_cast<int>(m),
                                         regressor.planePointers(), measurements.planePointers(),
                                         m_measurementCovariance.data(), m_lambda,
                                         state.planePointers(), covariance.planePointers(), stream);
    if (rc != BLF_CCM_OK)
        std::cerr << "[RecursiveLeastSquareBatch::advance] " << blf_ccm_last_error() << std::endl;
    return rc == BLF_CCM_OK;
}

bool RecursiveLeastSquareBatch::advanceContacts(const DeviceSoA& contactStates, const DeviceSoA* geometry,
                                                const DeviceSoA& measuredWrenches, DeviceSoA& state,
                                                DeviceSoA& covariance, void* stream)
{
    if (m_device == nullptr || contactStates.planes() != 30 || measuredWrenches.planes() != 6
        || state.planes() != 2 || covariance.planes() != 4 || m_measurementCovariance.size() != 6
        || (geometry != nullptr && geometry->planes() != 2))
    {
        std::cerr << "[RecursiveLeastSquareBatch::advanceContacts] Inconsistent plane counts."
                  << std::endl;
        return false;
    }
    const int rc = blf_ccm_rls_advance_contacts(raw(m_device),
                                                static_cast<std::int64_t>(contactStates.size()),
                                                contactStates.planePointers(),
                                                geometry ? geometry->planePointers() : nullptr,
                                                measuredWrenches.planePointers(),
                                                m_measurementCovariance.data(), m_lambda,
                                                state.planePointers(), covariance.planePointers(),
                                                stream);
    if (rc != BLF_CCM_OK)
        std::cerr << "[RecursiveLeastSquareBatch::advanceContacts] " << blf_ccm_last_error() << std::endl;
    return rc == BLF_CCM_OK;
}

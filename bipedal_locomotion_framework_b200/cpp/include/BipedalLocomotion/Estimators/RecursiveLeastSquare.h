/**
 * @file RecursiveLeastSquare.h
 * Recursive least squares (Ljung, System Identification, ch. 11.2) evaluated on a B200 through the
 * C ABI (blf_rls_advance_host / blf_rls_advance_batch / blf_ccm_rls_advance_contacts).
 *
 * RecursiveLeastSquare has the interface and state machine of the reference class
 * (src/Estimators/include/BipedalLocomotion/Estimators/RecursiveLeastSquare.h:28-111,
 * src/Estimators/src/RecursiveLeastSquare.cpp:17-149): parameters "measurement_covariance"
 * (vector), "lambda" (double), "state" (vector), "state_covariance" (vector).  Supported sizes:
 * 1..4 parameters, 1..6 measurements.  RecursiveLeastSquareBatch is the addition: n independent
 * estimators per call, and the fused contact-model identification step.
 */
#ifndef BIPEDAL_LOCOMOTION_ESTIMATORS_RLS_H
#define BIPEDAL_LOCOMOTION_ESTIMATORS_RLS_H

#include <functional>
#include <memory>
#include <vector>

#include <iDynTree/Core/MatrixDynSize.h>
#include <iDynTree/Core/VectorDynSize.h>

#include <BipedalLocomotion/ContactModels/ContinuousContactModelBatch.h>
#include <BipedalLocomotion/GenericContainer/DeviceSoA.h>
#include <BipedalLocomotion/ParametersHandler/IParametersHandler.h>

namespace BipedalLocomotion
{
namespace Estimators
{

class RecursiveLeastSquare
{
    iDynTree::VectorDynSize m_state;
    iDynTree::VectorDynSize m_measurements;
    iDynTree::MatrixDynSize m_stateCovarianceMatrix;
    std::vector<double> m_measurementCovariance; /**< diagonal: uncorrelated measurements */
    double m_lambda{1};
    std::function<iDynTree::MatrixDynSize(void)> m_regressor;

    enum class State
    {
        NotInitialized,
        Initialized,
        Running
    };
    State m_estimatorState{State::NotInitialized};
    std::shared_ptr<ContactModels::CudaDevice> m_device;

public:
    bool initialize(std::weak_ptr<ParametersHandler::IParametersHandler> handlerWeak);
    void setRegressorFunction(std::function<iDynTree::MatrixDynSize(void)> regressor);
    void setMeasurements(const iDynTree::VectorDynSize& measurements);
    /** One step of the filter (one n = 1 evaluation on the GPU). */
    bool advance();
    const iDynTree::VectorDynSize& parametersExpectedValue() const;
    const iDynTree::MatrixDynSize& parametersCovarianceMatrix() const;
};

/** n independent estimators resident on the device (structure-of-arrays planes). */
class RecursiveLeastSquareBatch
{
    std::shared_ptr<ContactModels::CudaDevice> m_device;
    std::vector<double> m_measurementCovariance;
    double m_lambda{1};

public:
    RecursiveLeastSquareBatch(std::shared_ptr<ContactModels::CudaDevice> device,
                              std::vector<double> measurementCovariance, double lambda);

    /** regressor: m*p planes (row-major m x p), measurements: m planes, state: p planes (in/out),
     * covariance: p*p planes (in/out). */
    bool advance(const GenericContainer::DeviceSoA& regressor,
                 const GenericContainer::DeviceSoA& measurements, GenericContainer::DeviceSoA& state,
                 GenericContainer::DeviceSoA& covariance, void* stream = nullptr);

    /** Contact-model identification: contactStates = the 30-plane state of
     * ContinuousContactModelBatch, measuredWrenches 6 planes, state = {spring, damper} planes,
     * covariance 4 planes; geometry = {length, width} planes or nullptr for the uniform geometry
     * set through ContinuousContactModelBatch::initialize on the same device. */
    bool advanceContacts(const GenericContainer::DeviceSoA& contactStates,
                         const GenericContainer::DeviceSoA* geometry,
                         const GenericContainer::DeviceSoA& measuredWrenches,
                         GenericContainer::DeviceSoA& state, GenericContainer::DeviceSoA& covariance,
                         void* stream = nullptr);
};

} // namespace Estimators
} // namespace BipedalLocomotion

#endif // BIPEDAL_LOCOMOTION_ESTIMATORS_RLS_H

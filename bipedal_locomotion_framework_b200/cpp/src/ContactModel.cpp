/**
 * @file ContactModel.cpp
 * Lazy-evaluation shell of the contact-model facade.
 *
 * Observable protocol as in the reference (src/ContactModels/src/ContactModel.cpp:12-92):
 * initialize() and both setters invalidate every cached result; a getter computes its result at
 * most once between two invalidations and otherwise serves the cached value -- which is also why
 * writing through springCoeff()/damperCoeff() leaves stale results in place, as upstream.
 * Kept as ONE validity mask whose bits are the C ABI's output mask, so that a derived class can
 * hand the same bit straight to blf_ccm_eval_batch_host.
 */
#include <BipedalLocomotion/ContactModels/ContactModel.h>

namespace BipedalLocomotion
{
namespace ContactModels
{
namespace
{
constexpr unsigned kWrench = 1u, kAutonomousDynamics = 2u, kControlMatrix = 4u, kRegressor = 8u;
}

void ContactModel::refresh(unsigned bit, void (ContactModel::*compute)())
{
    if (m_valid & bit) return;
    m_computeFailed = false;
    (this->*compute)();
    if (!m_computeFailed) m_valid |= bit;
}

bool ContactModel::initialize(std::weak_ptr<ParametersHandler::IParametersHandler> handler)
{
    m_valid = 0u;
    return initializePrivate(handler);
}

void ContactModel::setState(const iDynTree::Twist& twist, const iDynTree::Transform& transform)
{
    m_valid = 0u;
    setStatePrivate(twist, transform);
}

void ContactModel::setNullForceTransform(const iDynTree::Transform& transform)
{
    m_valid = 0u;
    setNullForceTransformPrivate(transform);
}

const iDynTree::Wrench& ContactModel::getContactWrench()
{
    refresh(kWrench, &ContactModel::computeContactWrench);
    return m_contactWrench;
}

const iDynTree::Vector6& ContactModel::getAutonomousDynamics()
{
    refresh(kAutonomousDynamics, &ContactModel::computeAutonomousDynamics);
    return m_autonomousDynamics;
}

const iDynTree::Matrix6x6& ContactModel::getControlMatrix()
{
    refresh(kControlMatrix, &ContactModel::computeControlMatrix);
    return m_controlMatrix;
}

const iDynTree::MatrixDynSize& ContactModel::getRegressor()
{
    refresh(kRegressor, &ContactModel::computeRegressor);
    return m_regressor;
}

} // namespace ContactModels
} // namespace BipedalLocomotion

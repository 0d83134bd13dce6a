/**
 * @file ContactModel.cpp
 * Lazy-evaluation shell of the contact-model facade.  Protocol as in the reference
 * (src/ContactModels/src/ContactModel.cpp:12-92): every setter and initialize() clear the four
 * "computed" flags; each getter computes once and then serves the cached value.
 */
#include <BipedalLocomotion/ContactModels/ContactModel.h>

using namespace BipedalLocomotion::ContactModels;

bool ContactModel::initialize(std::weak_ptr<ParametersHandler::IParametersHandler> handler)
{
    m_isContactWrenchComputed = false;
    m_isControlMatrixComputed = false;
    m_isAutonomousDynamicsComputed = false;
    m_isRegressorComputed = false;
    return initializePrivate(handler);
}

void ContactModel::setNullForceTransform(const iDynTree::Transform& nullForceTransform)
{
    m_isContactWrenchComputed = false;
    m_isControlMatrixComputed = false;
    m_isAutonomousDynamicsComputed = false;
    m_isRegressorComputed = false;
    setNullForceTransformPrivate(nullForceTransform);
}

void ContactModel::setState(const iDynTree::Twist& twist, const iDynTree::Transform& transform)
{
    m_isContactWrenchComputed = false;
    m_isControlMatrixComputed = false;
    m_isAutonomousDynamicsComputed = false;
    m_isRegressorComputed = false;
    setStatePrivate(twist, transform);
}

const iDynTree::Wrench& ContactModel::getContactWrench()
{
    if (!m_isContactWrenchComputed)
    {
        computeContactWrench();
        m_isContactWrenchComputed = true;
    }
    return m_contactWrench;
}

const iDynTree::Vector6& ContactModel::getAutonomousDynamics()
{
    if (!m_isAutonomousDynamicsComputed)
    {
        computeAutonomousDynamics();
        m_isAutonomousDynamicsComputed = true;
    }
    return m_autonomousDynamics;
}

const iDynTree::Matrix6x6& ContactModel::getControlMatrix()
{
    if (!m_isControlMatrixComputed)
    {
        computeControlMatrix();
        m_isControlMatrixComputed = true;
    }
    return m_controlMatrix;
}

const iDynTree::MatrixDynSize& ContactModel::getRegressor()
{
    if (!m_isRegressorComputed)
    {
        computeRegressor();
        m_isRegressorComputed = true;
    }
    return m_regressor;
}

/**
 * @file ContactModel.h
 * Abstract lazily-evaluated contact model: the per-instance facade of the B200 build.
 *
 * Public interface identical to the reference class
 * (src/ContactModels/include/BipedalLocomotion/ContactModels/ContactModel.h:33-146): initialize,
 * setState, setNullForceTransform and the four lazy getters that return const references to
 * internal storage valid until the next setter.  Same observable protocol as the four "computed"
 * flags of src/ContactModels/src/ContactModel.cpp:12-92, kept here as one validity mask.
 */
#ifndef BIPEDAL_LOCOMOTION_CONTACT_MODELS_CONTACT_MODEL_H
#define BIPEDAL_LOCOMOTION_CONTACT_MODELS_CONTACT_MODEL_H

#include <memory>

#include <iDynTree/Core/MatrixDynSize.h>
#include <iDynTree/Core/MatrixFixSize.h>
#include <iDynTree/Core/Transform.h>
#include <iDynTree/Core/Twist.h>
#include <iDynTree/Core/VectorFixSize.h>
#include <iDynTree/Core/Wrench.h>

#include <BipedalLocomotion/ParametersHandler/IParametersHandler.h>

namespace BipedalLocomotion
{
namespace ContactModels
{

// The C ABI (include/blf_ccm.h) takes arrays of these objects as flat doubles.
static_assert(sizeof(iDynTree::Twist) == 48, "Twist must be 6 packed doubles");
static_assert(sizeof(iDynTree::Wrench) == 48, "Wrench must be 6 packed doubles");
static_assert(sizeof(iDynTree::Vector6) == 48, "Vector6 must be 6 packed doubles");
static_assert(sizeof(iDynTree::Transform) == 96, "Transform must be position(3)+rotation(9)");
static_assert(sizeof(iDynTree::Matrix6x6) == 288, "Matrix6x6 must be 36 packed doubles");

class ContactModel
{
    /** Which cached results are current: one bit per getter, the bits of the C ABI's out_mask
     * (BLF_CCM_OUT_WRENCH = 1, _AUTODYN = 2, _CTRL = 4, _REGRESSOR = 8).  Cleared by every setter. */
    unsigned m_valid{0u};

    /** Run `compute` unless the result guarded by `bit` is current. */
    void refresh(unsigned bit, void (ContactModel::*compute)());

protected:
    /** A compute*() that could not produce its result (no device, failed launch) sets this; the
     * result is then NOT marked current, so the next getter tries again instead of serving the
     * untouched storage as if it were computed. */
    bool m_computeFailed{false};

    iDynTree::Wrench m_contactWrench; /**< contact wrench, mixed representation */
    iDynTree::Vector6 m_autonomousDynamics; /**< f of  d(wrench)/dt = f + g u */
    iDynTree::Matrix6x6 m_controlMatrix; /**< g of  d(wrench)/dt = f + g u */
    iDynTree::MatrixDynSize m_regressor; /**< wrench = regressor * [spring; damper] */

    virtual void computeContactWrench() = 0;
    virtual void computeAutonomousDynamics() = 0;
    virtual void computeControlMatrix() = 0;
    virtual void computeRegressor() = 0;
    virtual bool initializePrivate(std::weak_ptr<ParametersHandler::IParametersHandler> handler) = 0;
    virtual void setStatePrivate(const iDynTree::Twist& twist, const iDynTree::Transform& transform) = 0;
    virtual void setNullForceTransformPrivate(const iDynTree::Transform& transform) = 0;

public:
    virtual ~ContactModel() = default;

    /** Call before any other method.  The handler is only locked inside this call. */
    bool initialize(std::weak_ptr<ParametersHandler::IParametersHandler> handler);

    /** Get and compute (only if necessary) the contact wrench (mixed representation). */
    const iDynTree::Wrench& getContactWrench();
    /** Get and compute (only if necessary) the autonomous dynamics f. */
    const iDynTree::Vector6& getAutonomousDynamics();
    /** Get and compute (only if necessary) the control matrix g. */
    const iDynTree::Matrix6x6& getControlMatrix();
    /** Get and compute (only if necessary) the 6x2 regressor. */
    const iDynTree::MatrixDynSize& getRegressor();

    void setState(const iDynTree::Twist& twist, const iDynTree::Transform& transform);
    void setNullForceTransform(const iDynTree::Transform& transform);
};

} // namespace ContactModels
} // namespace BipedalLocomotion

#endif // BIPEDAL_LOCOMOTION_CONTACT_MODELS_CONTACT_MODEL_H

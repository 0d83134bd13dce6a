/**
 * @file ContactWrench.h
 * How the reference's dynamical systems hold a contact: the index of the contact frame plus shared
 * ownership of its contact model, handed out as a weak pointer
 * (src/System/include/BipedalLocomotion/System/ContactWrench.h:24-58, src/System/src/ContactWrench.cpp:13-35).
 * Same interface; a std::vector<ContactWrench> is what FloatingBaseDynamicalSystem's control input
 * carries and what ContactRolloutBatch::generalizedForce replaces the loop over.
 */
#ifndef BIPEDAL_LOCOMOTION_SYSTEM_CONTACT_WRENCH_H
#define BIPEDAL_LOCOMOTION_SYSTEM_CONTACT_WRENCH_H

#include <memory>

#include <iDynTree/Core/Wrench.h>
#include <iDynTree/Model/Indices.h>

#include <BipedalLocomotion/ContactModels/ContactModel.h>

namespace BipedalLocomotion
{
namespace System
{

class ContactWrench
{
    iDynTree::FrameIndex m_frame; /**< identifies the contact frame in the model */
    std::shared_ptr<ContactModels::ContactModel> m_contactModel;

public:
    ContactWrench(const iDynTree::FrameIndex& index, std::shared_ptr<ContactModels::ContactModel> model)
        : m_frame(index), m_contactModel(std::move(model))
    {
    }

    iDynTree::FrameIndex& index() noexcept { return m_frame; }
    const iDynTree::FrameIndex& index() const noexcept { return m_frame; }

    /** The holder keeps the model alive; callers lock the weak pointer for the duration of a use. */
    const std::weak_ptr<ContactModels::ContactModel> contactModel() const noexcept { return m_contactModel; }
};

} // namespace System
} // namespace BipedalLocomotion

#endif // BIPEDAL_LOCOMOTION_SYSTEM_CONTACT_WRENCH_H

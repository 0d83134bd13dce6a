// FreeFloatingState.h -- iDynTree::FreeFloatingGeneralizedTorques as FloatingBaseDynamicalSystem uses
// it (src/System/src/FloatingBaseSystemDynamics.cpp:68, :183-196): a base wrench plus one torque per
// internal degree of freedom, sized from the model.  Stand-in for the absent third-party header; with
// the real iDynTree (BLF_HAVE_IDYNTREE) include its own instead.
#ifndef BLF_IDYNTREE_MODEL_FREE_FLOATING_STATE_SHIM_H
#define BLF_IDYNTREE_MODEL_FREE_FLOATING_STATE_SHIM_H

#include <iDynTree/Core/CoreTypes.h>
#include <iDynTree/Model/Model.h>

namespace iDynTree
{

class FreeFloatingGeneralizedTorques
{
    Wrench m_base;
    VectorDynSize m_joints;

public:
    FreeFloatingGeneralizedTorques() = default;
    explicit FreeFloatingGeneralizedTorques(const Model& model) { resize(model); }
    void resize(const Model& model)
    {
        m_base.zero();
        m_joints.resize(model.getNrOfDOFs());
        m_joints.zero();
    }
    Wrench& baseWrench() { return m_base; }
    const Wrench& baseWrench() const { return m_base; }
    VectorDynSize& jointTorques() { return m_joints; }
    const VectorDynSize& jointTorques() const { return m_joints; }
};

} // namespace iDynTree

#endif // BLF_IDYNTREE_MODEL_FREE_FLOATING_STATE_SHIM_H

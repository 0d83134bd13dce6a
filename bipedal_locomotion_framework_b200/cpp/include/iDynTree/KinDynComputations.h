// KinDynComputations.h -- the part of iDynTree::KinDynComputations that
// FloatingBaseDynamicalSystem::dynamics calls (src/System/src/FloatingBaseSystemDynamics.cpp:64,
// :165-225): set the robot state, then ask for the mass matrix, the generalized bias forces and, per
// contact frame, the Jacobian, the frame velocity and the world transform.
//
// iDynTree's rigid-body algorithms are a third-party dependency that is absent here and out of this
// framework's scope: this header is an INTERFACE, not an implementation.  Whoever
// owns the robot model derives from it and answers the seven questions (with the real iDynTree: define
// BLF_HAVE_IDYNTREE, include its own header instead, and FloatingBaseDynamicalSystem takes the real
// class unchanged -- it calls nothing but these members).
#ifndef BLF_IDYNTREE_KINDYNCOMPUTATIONS_SHIM_H
#define BLF_IDYNTREE_KINDYNCOMPUTATIONS_SHIM_H

#include <iDynTree/Core/CoreTypes.h>
#include <iDynTree/Model/FreeFloatingState.h>
#include <iDynTree/Model/Indices.h>
#include <iDynTree/Model/Model.h>

namespace iDynTree
{

class KinDynComputations
{
public:
    virtual ~KinDynComputations() = default;

    virtual const Model& model() const = 0;

    /** World-to-base transform, joint positions, base twist (mixed representation), joint
     * velocities, gravity: the state every later answer refers to. */
    virtual bool setRobotState(const Transform& world_T_base, const VectorDynSize& jointPositions,
                               const Twist& baseVelocity, const VectorDynSize& jointVelocities,
                               const Vector3& worldGravity)
        = 0;

    /** (6 + dofs) x (6 + dofs), row-major. */
    virtual bool getFreeFloatingMassMatrix(MatrixDynSize& massMatrix) = 0;
    /** Coriolis + gravity terms: base wrench and joint torques. */
    virtual bool generalizedBiasForces(FreeFloatingGeneralizedTorques& biasForces) = 0;
    /** 6 x (6 + dofs), row-major. */
    virtual bool getFrameFreeFloatingJacobian(const FrameIndex frame, MatrixDynSize& jacobian) = 0;
    virtual Twist getFrameVel(const FrameIndex frame) = 0;
    virtual Transform getWorldTransform(const FrameIndex frame) = 0;
};

} // namespace iDynTree

#endif // BLF_IDYNTREE_KINDYNCOMPUTATIONS_SHIM_H

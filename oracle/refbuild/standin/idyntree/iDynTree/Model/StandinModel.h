// Stand-ins for iDynTree::Model, FreeFloatingGeneralizedTorques and -- as a TEST DOUBLE, not a
// kinematics library -- iDynTree::KinDynComputations.
//
// TEST INFRASTRUCTURE ONLY (see oracle/refbuild/README.md).  Purpose: compile the reference's
// src/System/src/FloatingBaseSystemDynamics.cpp UNMODIFIED, so that the contact loop of
// FloatingBaseDynamicalSystem::dynamics (:199-226, `m_knownCoefficent += J^T * wrench`) -- SURVEY
// section 8(f) row 2 -- can be run from the reference's own source.
//
// KinDynComputations here computes NOTHING: every quantity the reference asks it for (mass matrix,
// generalized bias forces, per-frame Jacobian, per-frame velocity and world transform) is a value
// the test INJECTED beforehand through the standin* setters.  setRobotState() only records that it
// was called.  What is exercised is therefore the reference's own sequencing and arithmetic around
// those quantities, nothing of iDynTree's rigid-body algorithms.
#ifndef BLF_REFBUILD_STANDIN_IDYNTREE_MODEL
#define BLF_REFBUILD_STANDIN_IDYNTREE_MODEL

#include <cstddef>
#include <map>
#include <string>

#include <iDynTree/Core/StandinCore.h>
#include <iDynTree/Model/Indices.h>

namespace iDynTree
{

class Model
{
    std::size_t m_dofs = 0;

public:
    Model() = default;
    explicit Model(std::size_t dofs) : m_dofs(dofs) {}
    std::size_t getNrOfDOFs() const { return m_dofs; }
    std::size_t getNrOfPosCoords() const { return m_dofs; }
    LinkIndex getFrameLink(FrameIndex frame) const { return frame; }
    std::string getFrameName(FrameIndex frame) const { return "frame" + std::to_string(frame); }
};

class JointDOFsDoubleArray : public VectorDynSize
{
public:
    JointDOFsDoubleArray() = default;
    explicit JointDOFsDoubleArray(std::size_t n) : VectorDynSize(n) {}
    explicit JointDOFsDoubleArray(const Model& m) : VectorDynSize(m.getNrOfDOFs()) {}
    void resize(std::size_t n) { VectorDynSize::resize(n); }
    void resize(const Model& m) { VectorDynSize::resize(m.getNrOfDOFs()); }
};

class FreeFloatingGeneralizedTorques
{
    Wrench m_baseWrench;
    JointDOFsDoubleArray m_jointTorques;

public:
    FreeFloatingGeneralizedTorques() { m_baseWrench.zero(); }
    explicit FreeFloatingGeneralizedTorques(const Model& m) { resize(m); }
    void resize(const Model& m)
    {
        m_baseWrench.zero();
        m_jointTorques.resize(m);
        m_jointTorques.zero();
    }
    Wrench& baseWrench() { return m_baseWrench; }
    JointDOFsDoubleArray& jointTorques() { return m_jointTorques; }
    const Wrench& baseWrench() const { return m_baseWrench; }
    const JointDOFsDoubleArray& jointTorques() const { return m_jointTorques; }
};

/// TEST DOUBLE: returns what was injected (see the header comment).
class KinDynComputations
{
    Model m_model;
    MatrixDynSize m_massMatrix;
    FreeFloatingGeneralizedTorques m_bias;
    struct Frame
    {
        MatrixDynSize jacobian;
        Twist velocity;
        Transform transform;
    };
    std::map<FrameIndex, Frame> m_frames;
    std::size_t m_setRobotStateCalls = 0;

public:
    // ---- injection (stand-in only) ----
    void standinSetModel(std::size_t dofs)
    {
        m_model = Model(dofs);
        m_bias.resize(m_model);
    }
    void standinSetMassMatrix(const MatrixDynSize& m) { m_massMatrix = m; }
    void standinSetBiasForces(const FreeFloatingGeneralizedTorques& h) { m_bias = h; }
    void standinSetFrame(FrameIndex frame, const MatrixDynSize& jacobian, const Twist& velocity,
                         const Transform& transform)
    {
        m_frames[frame] = Frame{jacobian, velocity, transform};
    }
    std::size_t standinSetRobotStateCalls() const { return m_setRobotStateCalls; }

    // ---- the part of iDynTree's interface the reference calls ----
    const Model& model() const { return m_model; }
    const Model& getRobotModel() const { return m_model; }
    bool setRobotState(const Transform&, const VectorDynSize& s, const Twist&, const VectorDynSize& sDot,
                       const Vector3&)
    {
        ++m_setRobotStateCalls;
        return s.size() == m_model.getNrOfDOFs() && sDot.size() == m_model.getNrOfDOFs();
    }
    bool getFreeFloatingMassMatrix(MatrixDynSize& out) const
    {
        if (m_massMatrix.rows() != 6 + m_model.getNrOfDOFs()) return false;
        out = m_massMatrix;
        return true;
    }
    bool generalizedBiasForces(FreeFloatingGeneralizedTorques& out) const
    {
        out = m_bias;
        return true;
    }
    bool getFrameFreeFloatingJacobian(FrameIndex frame, MatrixDynSize& out) const
    {
        auto it = m_frames.find(frame);
        if (it == m_frames.end()) return false;
        out = it->second.jacobian;
        return true;
    }
    Twist getFrameVel(FrameIndex frame) const { return m_frames.at(frame).velocity; }
    Transform getWorldTransform(FrameIndex frame) const { return m_frames.at(frame).transform; }
};

} // namespace iDynTree

#endif

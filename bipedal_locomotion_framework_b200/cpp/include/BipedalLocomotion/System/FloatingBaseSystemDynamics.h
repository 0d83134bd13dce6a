/**
 * @file FloatingBaseSystemDynamics.h
 * FloatingBaseDynamicalSystem, same interface as the reference's
 * src/System/include/BipedalLocomotion/System/FloatingBaseSystemDynamics.h:50-146 -- the class whose
 * dynamics() is the caller of the contact-model path (src/System/src/FloatingBaseSystemDynamics.cpp:199-226):
 *   state      = (base velocity (6, mixed), joint velocities, base position, base rotation, joint positions)
 *   derivative = (base acceleration, joint accelerations, base linear velocity, rotation rate, joint velocities)
 *   input      = (joint torques, contact wrenches)
 * The rigid-body quantities (mass matrix, bias forces, frame Jacobians / velocities / transforms) come
 * from the iDynTree::KinDynComputations object the caller sets, exactly as in the reference.  Everything
 * after them -- the contact wrenches of all contacts, -h + sum J^T wrench + torques, the
 * (M [+ regularization]).llt().solve, the rotation rate -- runs on the GPU through the C ABI
 * (blf_sys_floating_base_acceleration with one system: ONE upload, two launches, one download per
 * call, instead of one launch per contact model; blf_sys_kinematics_dynamics_host).  There is no CPU
 * evaluation path: without a device dynamics() returns false.  Many systems at once:
 * System::ContactRolloutBatch::floatingBaseAcceleration / floatingBaseEulerStep.
 */
#ifndef BIPEDAL_LOCOMOTION_SYSTEM_FLOATING_BASE_SYSTEM_DYNAMICS_H
#define BIPEDAL_LOCOMOTION_SYSTEM_FLOATING_BASE_SYSTEM_DYNAMICS_H

#include <memory>
#include <tuple>
#include <vector>

#include <iDynTree/Core/MatrixDynSize.h>
#include <iDynTree/Core/VectorFixSize.h>
#include <iDynTree/KinDynComputations.h>
#include <iDynTree/Model/FreeFloatingState.h>

#include <BipedalLocomotion/GenericContainer/DeviceSoA.h>
#include <BipedalLocomotion/ParametersHandler/IParametersHandler.h>
#include <BipedalLocomotion/System/ContactWrench.h>
#include <BipedalLocomotion/System/DynamicalSystem.h>
#include <BipedalLocomotion/System/StateTypes.h>

namespace BipedalLocomotion
{
namespace ContactModels
{
class CudaDevice;
}

namespace System
{

class FloatingBaseDynamicalSystem
    : public DynamicalSystem<std::tuple<Vector6d, VectorXd, Vector3d, Matrix3d, VectorXd>,
                             std::tuple<Vector6d, VectorXd, Vector3d, Matrix3d, VectorXd>,
                             std::tuple<VectorXd, std::vector<ContactWrench>>>
{
    static constexpr std::size_t m_baseDoFs = 6;

    std::shared_ptr<iDynTree::KinDynComputations> m_kinDyn;
    std::size_t m_actuatedDoFs{0};
    iDynTree::Vector3 m_gravity;

    iDynTree::MatrixDynSize m_massMatrix;
    iDynTree::FreeFloatingGeneralizedTorques m_generalizedBiasForces;
    iDynTree::MatrixDynSize m_jacobianMatrix;

    bool m_useMassMatrixRegularizationTerm{false};
    std::vector<double> m_massMatrixReglarizationTerm; /**< row-major, (6 + dofs)^2 */

    double m_rho{0.01}; /**< Baumgarte stabilization over SO(3) (reference default, :88) */

    // one staging block per call: everything dynamics() sends to the device, in one run of doubles
    std::shared_ptr<ContactModels::CudaDevice> m_device;
    int m_deviceIndex{-1};
    std::vector<double> m_staging;
    GenericContainer::DeviceSoA m_deviceBlock;
    std::vector<double> m_acceleration;

    bool ensureDevice(const char* where);

public:
    /** Gravity (0, 0, -9.81) as the reference's constructor; setGravityVector changes it. */
    FloatingBaseDynamicalSystem();
    explicit FloatingBaseDynamicalSystem(int device);
    explicit FloatingBaseDynamicalSystem(std::shared_ptr<ContactModels::CudaDevice> device);

    /** Reads the double parameter "rho" (FloatingBaseSystemDynamics.cpp:17-37). */
    bool initalize(std::weak_ptr<ParametersHandler::IParametersHandler> handler) override;

    void setGravityVector(const Vector3d& gravity);

    /** Sizes the buffers from kinDyn->model().getNrOfDOFs() (:53-74). */
    bool setKinDyn(std::shared_ptr<iDynTree::KinDynComputations> kinDyn);

    /** M + regularization is what gets factorised (:76-100).  rows x cols doubles, row-major (a
     * symmetric term reads the same either way); must be (6 + dofs) square. */
    bool setMassMatrixRegularization(const double* matrix, std::size_t rows, std::size_t cols);
    bool setMassMatrixRegularization(const iDynTree::MatrixDynSize& matrix);

    bool dynamics(const double& time, StateDerivativeType& stateDerivative) final;

    ~FloatingBaseDynamicalSystem() = default;
};

} // namespace System
} // namespace BipedalLocomotion

#endif // BIPEDAL_LOCOMOTION_SYSTEM_FLOATING_BASE_SYSTEM_DYNAMICS_H

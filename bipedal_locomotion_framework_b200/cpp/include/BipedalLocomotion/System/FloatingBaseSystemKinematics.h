/**
 * @file FloatingBaseSystemKinematics.h
 * Kinematics of a floating-base system, same interface as the reference's
 * src/System/include/BipedalLocomotion/System/FloatingBaseSystemKinematics.h:25-58:
 *   state      = (base position, base rotation, joint positions)
 *   derivative = (base linear velocity, rotation rate, joint velocities)
 *   input      = (base twist in mixed representation, joint velocities)
 * dynamics() (FloatingBaseSystemKinematics.cpp:36-73) and the Euler update run on the GPU through
 * the C ABI (blf_sys_kinematics_dynamics_host / blf_sys_kinematics_integrate_host); there is no CPU
 * evaluation path: without a device both return false.
 */
#ifndef BIPEDAL_LOCOMOTION_SYSTEM_FLOATING_BASE_SYSTEM_KINEMATICS_H
#define BIPEDAL_LOCOMOTION_SYSTEM_FLOATING_BASE_SYSTEM_KINEMATICS_H

#include <memory>
#include <tuple>

#include <BipedalLocomotion/ParametersHandler/IParametersHandler.h>
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

class FloatingBaseSystemKinematics
    : public DynamicalSystem<std::tuple<Vector3d, Matrix3d, VectorXd>,
                             std::tuple<Vector3d, Matrix3d, VectorXd>,
                             std::tuple<Vector6d, VectorXd>>
{
    double m_rho{0.01}; /**< Baumgarte stabilization over SO(3) (reference default, :36) */
    std::shared_ptr<ContactModels::CudaDevice> m_device;
    int m_deviceIndex{0};

    bool ensureDevice(const char* where);

public:
    FloatingBaseSystemKinematics() = default;
    explicit FloatingBaseSystemKinematics(int device) : m_deviceIndex(device) {}
    explicit FloatingBaseSystemKinematics(std::shared_ptr<ContactModels::CudaDevice> device)
        : m_device(std::move(device))
    {
    }

    /** Reads the double parameter "rho" (:13-34). */
    bool initalize(std::weak_ptr<ParametersHandler::IParametersHandler> handler) override;

    bool dynamics(const double& time, StateDerivativeType& stateDerivative) final;

    /** steps-1 Euler steps of stepDT and one of lastDT with the current control input, state
     * updated in place -- the hook ForwardEuler uses (Integrator.h). */
    bool advanceOnDevice(double stepDT, double lastDT, int steps);

    ~FloatingBaseSystemKinematics() = default;
};

} // namespace System
} // namespace BipedalLocomotion

#endif // BIPEDAL_LOCOMOTION_SYSTEM_FLOATING_BASE_SYSTEM_KINEMATICS_H

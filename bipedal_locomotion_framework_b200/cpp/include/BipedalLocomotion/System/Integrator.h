/**
 * @file Integrator.h
 * Integrator / FixedStepIntegrator / ForwardEuler with the reference's interface
 * (src/System/include/BipedalLocomotion/System/Integrator.h:27-79, FixedStepIntegrator.h:24-58,
 * ForwardEuler.h:30-69 and the matching .tpp files).
 *
 * FixedStepIntegrator::integrate reproduces the reference's step schedule exactly
 * (FixedStepIntegrator.tpp:48-64): iterations = ceil((tf - t0) / dT) steps, the first
 * iterations-1 of size dT, the last of size tf - currentTime where currentTime was only advanced
 * inside the loop (so with >= 2 iterations the last step is roughly 2 dT -- kept, not fixed).
 * One deviation: tf == t0 gives iterations == 0, for which the reference's loop bound underflows
 * and never terminates; here integrate() returns false.
 *
 * Device hook: a system that defines
 *     bool advanceOnDevice(double stepDT, double lastDT, int steps);
 * (FloatingBaseSystemKinematics does) is advanced by ForwardEuler with ONE call for the whole
 * schedule -- dynamics and x += dx*dT both run on the GPU through the C ABI.  Any other system goes
 * through the generic host path below (dynamics() + tuple-wise x += dx*dT), as in the reference.
 */
#ifndef BIPEDAL_LOCOMOTION_SYSTEM_INTEGRATOR_H
#define BIPEDAL_LOCOMOTION_SYSTEM_INTEGRATOR_H

#include <cmath>
#include <cstddef>
#include <iostream>
#include <memory>
#include <tuple>
#include <type_traits>
#include <utility>

#include <BipedalLocomotion/System/DynamicalSystem.h>

namespace BipedalLocomotion
{
namespace System
{

template <typename DynamicalSystemDerived> class Integrator
{
    static_assert(std::is_base_of<DynamicalSystem<typename DynamicalSystemDerived::StateType,
                                                  typename DynamicalSystemDerived::StateDerivativeType,
                                                  typename DynamicalSystemDerived::InputType>,
                                  DynamicalSystemDerived>::value,
                  "The integrator template type has to be derived from DynamicalSystem.");

protected:
    std::shared_ptr<DynamicalSystemDerived> m_dynamicalSystem;

public:
    /** The dynamical system can be set only once. */
    bool setDynamicalSystem(std::shared_ptr<DynamicalSystemDerived> dynamicalSystem)
    {
        if (m_dynamicalSystem != nullptr)
        {
            std::cerr << "[Integrator::setDynamicalSystem] The dynamical system has been already set."
                      << std::endl;
            return false;
        }
        if (dynamicalSystem == nullptr)
        {
            std::cerr << "[Integrator::setDynamicalSystem] The dynamical system passed to the function "
                         "is corrupted."
                      << std::endl;
            return false;
        }
        m_dynamicalSystem = dynamicalSystem;
        return true;
    }

    const std::weak_ptr<DynamicalSystemDerived> dynamicalSystem() const { return m_dynamicalSystem; }

    const typename DynamicalSystemDerived::StateType& getSolution() const
    {
        return m_dynamicalSystem->getState();
    }

    virtual bool integrate(double initialTime, double finalTime) = 0;

    ~Integrator() = default;
};

template <typename DynamicalSystemDerived>
class FixedStepIntegrator : public Integrator<DynamicalSystemDerived>
{
protected:
    double m_dT{0.0};

    virtual bool oneStepIntegration(double t0, double dT) = 0;

    /** `iterations` steps: iterations-1 of m_dT starting at t0, then one of lastDT.  The default
     * walks them one by one; ForwardEuler replaces it for systems with a device hook. */
    virtual bool integrateSchedule(double initialTime, int iterations, double lastStepTime, double lastDT)
    {
        for (int i = 0; i < iterations - 1; i++)
        {
            const double currentTime = initialTime + m_dT * i;
            if (!oneStepIntegration(currentTime, m_dT))
            {
                std::cerr << "[FixedStepIntegrator::integrate] Error while integrating at time: "
                          << currentTime << "." << std::endl;
                return false;
            }
        }
        if (!oneStepIntegration(lastStepTime, lastDT))
        {
            std::cerr << "[FixedStepIntegrator::integrate] Error while integrating the last step."
                      << std::endl;
            return false;
        }
        return true;
    }

public:
    FixedStepIntegrator(const double& dT) : m_dT{dT} {}

    bool integrate(double initialTime, double finalTime) final
    {
        if (this->m_dynamicalSystem == nullptr)
        {
            std::cerr << "[FixedStepIntegrator::integrate] Please set the dynamical system before call "
                         "this function."
                      << std::endl;
            return false;
        }
        if (initialTime > finalTime)
        {
            std::cerr << "[FixedStepIntegrator::integrate] The final time has to be greater than the "
                         "initial one."
                      << std::endl;
            return false;
        }
        if (m_dT <= 0)
        {
            std::cerr << "[FixedStepIntegrator::integrate] The sampling time must be a strictly "
                         "positive number."
                      << std::endl;
            return false;
        }
        const int iterations = static_cast<int>(std::ceil((finalTime - initialTime) / m_dT));
        if (iterations < 1)
        {
            std::cerr << "[FixedStepIntegrator::integrate] The final time is equal to the initial "
                         "one: nothing to integrate."
                      << std::endl;
            return false;
        }
        double currentTime = initialTime;
        for (int i = 0; i < iterations - 1; i++) currentTime = initialTime + m_dT * i;
        return integrateSchedule(initialTime, iterations, currentTime, finalTime - currentTime);
    }

    ~FixedStepIntegrator() = default;
};

namespace detail
{
template <typename T, typename = void> struct HasDeviceAdvance : std::false_type
{
};
template <typename T>
struct HasDeviceAdvance<T, std::void_t<decltype(std::declval<T&>().advanceOnDevice(0.0, 0.0, 1))>>
    : std::true_type
{
};
} // namespace detail

/**
 * Forward Euler integration method.
 * @warning operator+= and operator*(double) must exist for the objects contained in
 * StateType and StateDerivativeType (as in the reference, ForwardEuler.h:26-28).
 */
template <typename DynamicalSystemDerived> class ForwardEuler : public FixedStepIntegrator<DynamicalSystemDerived>
{
    typename DynamicalSystemDerived::StateDerivativeType m_computationalBufferStateDerivative;
    typename DynamicalSystemDerived::StateType m_computationalBufferState;

    template <std::size_t I = 0, typename... Tp, typename... Td>
    void addArea(const std::tuple<Tp...>& dx, const double& dT, std::tuple<Td...>& x)
    {
        static_assert(sizeof...(Tp) == sizeof...(Td));
        if constexpr (I < sizeof...(Tp))
        {
            std::get<I>(x) += std::get<I>(dx) * dT;
            addArea<I + 1>(dx, dT, x);
        }
    }

    bool oneStepIntegration(double t0, double dT) final
    {
        if (this->m_dynamicalSystem == nullptr)
        {
            std::cerr << "[ForwardEuler::oneStepIntegration] Please specify the dynamical system."
                      << std::endl;
            return false;
        }
        if constexpr (detail::HasDeviceAdvance<DynamicalSystemDerived>::value)
        {
            (void)t0;
            return this->m_dynamicalSystem->advanceOnDevice(dT, dT, 1);
        } else
        {
            if (!this->m_dynamicalSystem->dynamics(t0, m_computationalBufferStateDerivative))
            {
                std::cerr << "[ForwardEuler::oneStepIntegration] Unable to compute the system dynamics."
                          << std::endl;
                return false;
            }
            // x = x0 + dT * dx
            m_computationalBufferState = this->m_dynamicalSystem->getState();
            addArea(m_computationalBufferStateDerivative, dT, m_computationalBufferState);
            if (!this->m_dynamicalSystem->setState(m_computationalBufferState))
            {
                std::cerr << "[ForwardEuler::oneStepIntegration] Unable to set the new state in the "
                             "dynamical system."
                          << std::endl;
                return false;
            }
            return true;
        }
    }

    bool integrateSchedule(double initialTime, int iterations, double lastStepTime, double lastDT) final
    {
        if constexpr (detail::HasDeviceAdvance<DynamicalSystemDerived>::value)
        {
            (void)initialTime;
            (void)lastStepTime;
            if (!this->m_dynamicalSystem->advanceOnDevice(this->m_dT, lastDT, iterations))
            {
                std::cerr << "[FixedStepIntegrator::integrate] Error while integrating on the device."
                          << std::endl;
                return false;
            }
            return true;
        } else
        {
            return FixedStepIntegrator<DynamicalSystemDerived>::integrateSchedule(initialTime, iterations,
                                                                                  lastStepTime, lastDT);
        }
    }

public:
    ForwardEuler(const double& dT) : FixedStepIntegrator<DynamicalSystemDerived>(dT) {}
    ~ForwardEuler() = default;
};

} // namespace System
} // namespace BipedalLocomotion

#endif // BIPEDAL_LOCOMOTION_SYSTEM_INTEGRATOR_H

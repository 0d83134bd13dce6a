/**
 * @file DynamicalSystem.h
 * Base of every dynamical system, same interface as the reference's
 * src/System/include/BipedalLocomotion/System/DynamicalSystem.h:33-100 (+ DynamicalSystem.tpp):
 * State, StateDerivative and Input are std::tuple specialisations; setState / setControlInput
 * store, getState returns a const reference, dynamics() is the derived class's job.
 */
#ifndef BIPEDAL_LOCOMOTION_SYSTEM_DYNAMICAL_SYSTEM_H
#define BIPEDAL_LOCOMOTION_SYSTEM_DYNAMICAL_SYSTEM_H

#include <memory>
#include <tuple>
#include <type_traits>

#include <BipedalLocomotion/ParametersHandler/IParametersHandler.h>

namespace BipedalLocomotion
{
namespace System
{

namespace detail
{
template <typename T> struct IsTuple : std::false_type
{
};
template <typename... Ts> struct IsTuple<std::tuple<Ts...>> : std::true_type
{
};
} // namespace detail

template <typename State, typename StateDerivative, typename Input> class DynamicalSystem
{
    static_assert(detail::IsTuple<State>::value, "The State type must be a specialization of the std::tuple.");
    static_assert(detail::IsTuple<StateDerivative>::value,
                  "The StateDerivative type must be a specialization of the std::tuple.");
    static_assert(detail::IsTuple<Input>::value, "The Input type must be a specialization of the std::tuple.");

public:
    using StateType = State;
    using StateDerivativeType = StateDerivative;
    using InputType = Input;

protected:
    InputType m_controlInput;
    StateType m_state;

public:
    /** [sic] the reference spells it `initalize` (DynamicalSystem.h:62). */
    virtual bool initalize(std::weak_ptr<ParametersHandler::IParametersHandler> /*handler*/) { return true; }

    virtual bool setState(const StateType& state)
    {
        m_state = state;
        return true;
    }

    const StateType& getState() const { return m_state; }

    virtual bool setControlInput(const InputType& controlInput)
    {
        m_controlInput = controlInput;
        return true;
    }

    virtual bool dynamics(const double& time, StateDerivativeType& stateDerivative) = 0;

    ~DynamicalSystem() = default;
};

} // namespace System
} // namespace BipedalLocomotion

#endif // BIPEDAL_LOCOMOTION_SYSTEM_DYNAMICAL_SYSTEM_H

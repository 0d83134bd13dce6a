/**
 * @file StdImplementation.h
 * std::unordered_map<std::string, std::any> backed parameters handler.
 *
 * Same behaviour as the reference's StdImplementation
 * (src/ParametersHandler/include/BipedalLocomotion/ParametersHandler/StdImplementation.h:27-236,
 * StdImplementation.tpp:21-105, src/StdImplementation.cpp:14-169): strict std::any_cast typing (an
 * int stored under a key cannot be read as double), bool + std::cerr errors, groups stored as
 * shared pointers, getGroup() of a missing name yields an expired weak pointer.
 */
#ifndef BIPEDAL_LOCOMOTION_PARAMETERS_HANDLER_STD_IMPLEMENTATION_H
#define BIPEDAL_LOCOMOTION_PARAMETERS_HANDLER_STD_IMPLEMENTATION_H

#include <any>
#include <iostream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include <BipedalLocomotion/ParametersHandler/IParametersHandler.h>

namespace BipedalLocomotion
{
namespace ParametersHandler
{

class StdImplementation : public IParametersHandler
{
    std::unordered_map<std::string, std::any> m_map;

    template <typename T> bool getParameterPrivate(const std::string& parameterName, T& parameter) const
    {
        auto it = m_map.find(parameterName);
        if (it == m_map.end())
        {
            std::cerr << "[StdImplementation::getParameterPrivate] Parameter named " << parameterName
                      << " not found." << std::endl;
            return false;
        }
        const T* value = std::any_cast<T>(&it->second);
        if (value == nullptr)
        {
            std::cerr << "[StdImplementation::getParameterPrivate] The type of the parameter named "
                      << parameterName << " is different from the one expected" << std::endl;
            return false;
        }
        parameter = *value;
        return true;
    }

    /** Vector parameters: the reference's resize contract (StdImplementation.tpp:62-85). */
    template <typename T>
    bool getVectorPrivate(const std::string& parameterName, std::vector<T>& parameter,
                          GenericContainer::VectorResizeMode mode) const
    {
        std::vector<T> stored;
        if (!getParameterPrivate(parameterName, stored)) return false;
        if (stored.size() != parameter.size() && mode != GenericContainer::VectorResizeMode::Resizable)
        {
            std::cerr << "[StdImplementation::getParameterPrivate] Unable to resize the vector. List size: "
                      << stored.size() << ". Vector size: " << parameter.size() << std::endl;
            return false;
        }
        parameter = stored;
        return true;
    }

    /** The same contract through a generic view: the view resizes its owner when it may. */
    template <typename T>
    bool getViewPrivate(const std::string& parameterName, GenericContainer::Vector<T>& parameter) const
    {
        std::vector<T> stored;
        if (!getParameterPrivate(parameterName, stored)) return false;
        const auto want = static_cast<typename GenericContainer::Vector<T>::index_type>(stored.size());
        if (parameter.size() != want && !parameter.resizeVector(want))
        {
            std::cerr << "[StdImplementation::getParameterPrivate] Unable to resize the vector. List size: "
                      << stored.size() << ". Vector size: " << parameter.size() << std::endl;
            return false;
        }
        for (std::size_t i = 0; i < stored.size(); ++i) parameter[static_cast<std::ptrdiff_t>(i)] = stored[i];
        return true;
    }

public:
    using IParametersHandler::getParameter;   // the container templates of the interface
    using IParametersHandler::setParameter;
    // unique_ptr / shared_ptr / weak_ptr are the ones inherited from IParametersHandler (pointers to
    // the INTERFACE), as upstream: `StdImplementation::shared_ptr g = h->getGroup("x").lock();`

    StdImplementation() = default;
    explicit StdImplementation(const std::unordered_map<std::string, std::any>& map) : m_map(map) {}

    bool getParameter(const std::string& n, int& p) const final { return getParameterPrivate(n, p); }
    bool getParameter(const std::string& n, double& p) const final { return getParameterPrivate(n, p); }
    bool getParameter(const std::string& n, std::string& p) const final { return getParameterPrivate(n, p); }
    bool getParameter(const std::string& n, bool& p) const final { return getParameterPrivate(n, p); }
    bool getParameter(const std::string& n, std::vector<bool>& p) const final { return getParameterPrivate(n, p); }
    bool getParameter(const std::string& n, std::vector<int>& p,
                      GenericContainer::VectorResizeMode mode = GenericContainer::VectorResizeMode::Fixed) const final
    {
        return getVectorPrivate(n, p, mode);
    }
    bool getParameter(const std::string& n, std::vector<double>& p,
                      GenericContainer::VectorResizeMode mode = GenericContainer::VectorResizeMode::Fixed) const final
    {
        return getVectorPrivate(n, p, mode);
    }
    bool getParameter(const std::string& n, std::vector<std::string>& p,
                      GenericContainer::VectorResizeMode mode = GenericContainer::VectorResizeMode::Fixed) const final
    {
        return getVectorPrivate(n, p, mode);
    }

    bool getParameter(const std::string& n, GenericContainer::Vector<int>& p) const final { return getViewPrivate(n, p); }
    bool getParameter(const std::string& n, GenericContainer::Vector<double>& p) const final { return getViewPrivate(n, p); }
    bool getParameter(const std::string& n, GenericContainer::Vector<std::string>& p) const final
    {
        return getViewPrivate(n, p);
    }

    void setParameter(const std::string& n, const int& p) final { m_map[n] = p; }
    void setParameter(const std::string& n, const double& p) final { m_map[n] = p; }
    void setParameter(const std::string& n, const std::string& p) final { m_map[n] = p; }
    void setParameter(const std::string& n, const char* p) final { m_map[n] = std::string(p); }
    void setParameter(const std::string& n, const bool& p) final { m_map[n] = p; }
    void setParameter(const std::string& n, const std::vector<bool>& p) final { m_map[n] = p; }
    void setParameter(const std::string& n, const std::vector<int>& p) final { m_map[n] = p; }
    void setParameter(const std::string& n, const std::vector<double>& p) final { m_map[n] = p; }
    void setParameter(const std::string& n, const std::vector<std::string>& p) final { m_map[n] = p; }

    IParametersHandler::weak_ptr getGroup(const std::string& name) const final
    {
        auto it = m_map.find(name);
        if (it == m_map.end()) return std::make_shared<StdImplementation>(); // expires at once
        const auto* group = std::any_cast<std::shared_ptr<StdImplementation>>(&it->second);
        if (group == nullptr)
        {
            std::cerr << "[StdImplementation::getGroup] The element named " << name
                      << " is not a group" << std::endl;
            return std::make_shared<StdImplementation>();
        }
        return *group;
    }

    bool setGroup(const std::string& name, IParametersHandler::shared_ptr newGroup) final
    {
        auto down = std::dynamic_pointer_cast<StdImplementation>(newGroup);
        if (down == nullptr)
        {
            std::cerr << "[StdImplementation::setGroup] Unable to downcast the pointer to "
                         "StdImplementation."
                      << std::endl;
            return false;
        }
        m_map[name] = std::make_any<std::shared_ptr<StdImplementation>>(down);
        return true;
    }

    void set(const std::unordered_map<std::string, std::any>& object) { m_map = object; }

    std::string toString() const final
    {
        std::string keys;
        for (const auto& kv : m_map) keys += kv.first + " ";
        return keys;
    }
    bool isEmpty() const final { return m_map.empty(); }
    void clear() final { m_map.clear(); }
    ~StdImplementation() = default;
};

} // namespace ParametersHandler
} // namespace BipedalLocomotion

#endif // BIPEDAL_LOCOMOTION_PARAMETERS_HANDLER_STD_IMPLEMENTATION_H

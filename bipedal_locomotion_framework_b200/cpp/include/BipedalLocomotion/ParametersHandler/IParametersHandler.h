/**
 * @file IParametersHandler.h
 * Parameters handler interface of the B200 contact-model build.
 *
 * Mirrors the public surface of the reference interface
 * (src/ParametersHandler/include/BipedalLocomotion/ParametersHandler/IParametersHandler.h:26-249):
 * same method names, argument meaning and bool-return error convention.  Vector parameters are
 * carried as std::vector<T> (the reference routes them through GenericContainer::Vector, a host
 * utility that is outside this build's scope) and keep the reference's resize contract: by default
 * (VectorResizeMode::Fixed) the destination must already have the size of the stored list, pass
 * VectorResizeMode::Resizable to have it resized (IParametersHandler.h:129-139,
 * StdImplementation.tpp:62-85).  std::vector<bool> is always resized, as upstream.
 */
#ifndef BIPEDAL_LOCOMOTION_PARAMETERS_HANDLER_IPARAMETERS_HANDLER_H
#define BIPEDAL_LOCOMOTION_PARAMETERS_HANDLER_IPARAMETERS_HANDLER_H

#include <memory>
#include <string>
#include <vector>

#include <BipedalLocomotion/GenericContainer/Vector.h>

namespace BipedalLocomotion
{
namespace ParametersHandler
{

class IParametersHandler
{
public:
    using unique_ptr = std::unique_ptr<IParametersHandler>;
    using shared_ptr = std::shared_ptr<IParametersHandler>;
    using weak_ptr = std::weak_ptr<IParametersHandler>;

    /** Get a parameter; false (and a message on std::cerr) if missing or of a different type. */
    virtual bool getParameter(const std::string& parameterName, int& parameter) const = 0;
    virtual bool getParameter(const std::string& parameterName, double& parameter) const = 0;
    virtual bool getParameter(const std::string& parameterName, std::string& parameter) const = 0;
    virtual bool getParameter(const std::string& parameterName, bool& parameter) const = 0;
    virtual bool getParameter(const std::string& parameterName, std::vector<bool>& parameter) const = 0;
    virtual bool getParameter(const std::string& parameterName, std::vector<int>& parameter,
                              GenericContainer::VectorResizeMode mode = GenericContainer::VectorResizeMode::Fixed) const = 0;
    virtual bool getParameter(const std::string& parameterName, std::vector<double>& parameter,
                              GenericContainer::VectorResizeMode mode = GenericContainer::VectorResizeMode::Fixed) const = 0;
    virtual bool getParameter(const std::string& parameterName, std::vector<std::string>& parameter,
                              GenericContainer::VectorResizeMode mode = GenericContainer::VectorResizeMode::Fixed) const = 0;

    virtual void setParameter(const std::string& parameterName, const int& parameter) = 0;
    virtual void setParameter(const std::string& parameterName, const double& parameter) = 0;
    virtual void setParameter(const std::string& parameterName, const std::string& parameter) = 0;
    /** Needed so that a string literal does not bind to the bool overload. */
    virtual void setParameter(const std::string& parameterName, const char* parameter) = 0;
    virtual void setParameter(const std::string& parameterName, const bool& parameter) = 0;
    virtual void setParameter(const std::string& parameterName, const std::vector<bool>& parameter) = 0;
    virtual void setParameter(const std::string& parameterName, const std::vector<int>& parameter) = 0;
    virtual void setParameter(const std::string& parameterName, const std::vector<double>& parameter) = 0;
    virtual void setParameter(const std::string& parameterName, const std::vector<std::string>& parameter) = 0;

    /** Get a group; if it does not exist the returned weak pointer cannot be locked. */
    virtual weak_ptr getGroup(const std::string& name) const = 0;
    virtual bool setGroup(const std::string& name, shared_ptr newGroup) = 0;
    virtual std::string toString() const = 0;
    virtual bool isEmpty() const = 0;
    virtual void clear() = 0;
    virtual ~IParametersHandler() = default;
};

} // namespace ParametersHandler
} // namespace BipedalLocomotion

#endif // BIPEDAL_LOCOMOTION_PARAMETERS_HANDLER_IPARAMETERS_HANDLER_H

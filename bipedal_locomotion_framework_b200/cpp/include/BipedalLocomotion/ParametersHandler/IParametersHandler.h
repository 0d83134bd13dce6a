/**
 * @file IParametersHandler.h
 * Parameters handler interface of the B200 contact-model build.
 *
 * Mirrors the public surface of the reference interface
 * (src/ParametersHandler/include/BipedalLocomotion/ParametersHandler/IParametersHandler.h:26-249):
 * same method names, argument meaning and bool-return error convention.  Vector parameters can be
 * read into / set from std::vector<T> directly and, as in the reference, from ANY container a
 * GenericContainer::Vector can view (std::array, iDynTree::VectorDynSize, Eigen vectors, plain
 * arrays: the templates below).  The reference's resize contract holds for all of them: by default
 * (VectorResizeMode::Fixed) the destination must already have the size of the stored list, pass
 * VectorResizeMode::Resizable to have it resized (IParametersHandler.h:129-139,
 * StdImplementation.tpp:62-85).  std::vector<bool> is always resized, as upstream.
 */
#ifndef BIPEDAL_LOCOMOTION_PARAMETERS_HANDLER_IPARAMETERS_HANDLER_H
#define BIPEDAL_LOCOMOTION_PARAMETERS_HANDLER_IPARAMETERS_HANDLER_H

#include <iterator>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#include <BipedalLocomotion/GenericContainer/Vector.h>

namespace BipedalLocomotion
{
namespace ParametersHandler
{

class IParametersHandler
{
    template <typename T> struct is_std_vector : std::false_type
    {
    };
    template <typename T, typename A> struct is_std_vector<std::vector<T, A>> : std::true_type
    {
    };

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

    /** Vector parameters through a generic view (any element storage; see the template below). */
    virtual bool getParameter(const std::string& parameterName, GenericContainer::Vector<int>& parameter) const = 0;
    virtual bool getParameter(const std::string& parameterName, GenericContainer::Vector<double>& parameter) const = 0;
    virtual bool getParameter(const std::string& parameterName, GenericContainer::Vector<std::string>& parameter) const = 0;

    /** Get a vector parameter into any container a GenericContainer::Vector can view (same
     * template as the reference's, IParametersHandler.h:121-139). */
    template <class Container,
              typename = std::enable_if_t<!GenericContainer::is_vector<Container>::value
                                          && !is_std_vector<Container>::value
                                          && !std::is_same<Container, std::string>::value
                                          && GenericContainer::is_vector_constructible<Container>::value>>
    bool getParameter(const std::string& parameterName, Container& parameter,
                      GenericContainer::VectorResizeMode mode = GenericContainer::VectorResizeMode::Fixed) const
    {
        auto view = GenericContainer::make_vector(parameter, mode);
        return this->getParameter(parameterName, view);
    }

    /** Set a vector parameter from any such container (stored as std::vector<T>). */
    template <class Container,
              typename = std::enable_if_t<!is_std_vector<Container>::value
                                          && !std::is_same<Container, std::string>::value
                                          && !std::is_convertible<Container, const char*>::value
                                          && (GenericContainer::is_vector<Container>::value
                                              || GenericContainer::is_vector_constructible<Container>::value)>>
    void setParameter(const std::string& parameterName, const Container& parameter)
    {
        using T = std::remove_cv_t<std::remove_reference_t<decltype(*std::begin(parameter))>>;
        this->setParameter(parameterName, std::vector<T>(std::begin(parameter), std::end(parameter)));
    }

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

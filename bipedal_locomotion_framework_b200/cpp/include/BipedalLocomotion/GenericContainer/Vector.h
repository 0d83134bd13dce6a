/**
 * @file Vector.h
 * The one piece of the reference's GenericContainer::Vector machinery that is visible in the
 * parameters-handler interface: the resize mode a caller passes to getParameter for vectors
 * (src/GenericContainer/include/BipedalLocomotion/GenericContainer/Vector.h, VectorResizeMode;
 * used at src/ParametersHandler/include/BipedalLocomotion/ParametersHandler/IParametersHandler.h:129-139).
 * The generic non-owning Vector<T> itself is a host utility outside this build's scope; vector
 * parameters travel as std::vector<T>.
 */
#ifndef BIPEDAL_LOCOMOTION_GENERIC_CONTAINER_VECTOR_H
#define BIPEDAL_LOCOMOTION_GENERIC_CONTAINER_VECTOR_H

namespace BipedalLocomotion
{
namespace GenericContainer
{
/** Fixed (the reference's default): the destination must already have the size of the stored
 * list, otherwise getParameter fails.  Resizable: the destination is resized to it. */
enum class VectorResizeMode
{
    Resizable,
    Fixed
};
} // namespace GenericContainer
} // namespace BipedalLocomotion

#endif // BIPEDAL_LOCOMOTION_GENERIC_CONTAINER_VECTOR_H

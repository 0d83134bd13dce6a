/**
 * @file TemplateHelpers.h
 * Present so that code including the reference's header of the same name compiles against this
 * build; the metaprogramming helpers of the reference
 * (src/GenericContainer/include/BipedalLocomotion/GenericContainer/TemplateHelpers.h) serve its
 * GenericContainer::Vector, which this build does not carry.
 */
#ifndef BIPEDAL_LOCOMOTION_TEMPLATEHELPERS_H
#define BIPEDAL_LOCOMOTION_TEMPLATEHELPERS_H
#include <BipedalLocomotion/GenericContainer/Vector.h>
#endif // BIPEDAL_LOCOMOTION_TEMPLATEHELPERS_H

/**
 * @file TemplateHelpers.h
 * Present so that code including the reference's header of the same name compiles against this
 * build: the traits GenericContainer::Vector needs (is_vector, is_vector_constructible and the
 * data()/size()/resize() detectors) live in Vector.h here.
 */
#ifndef BIPEDAL_LOCOMOTION_TEMPLATEHELPERS_H
#define BIPEDAL_LOCOMOTION_TEMPLATEHELPERS_H
#include <BipedalLocomotion/GenericContainer/Vector.h>
#endif // BIPEDAL_LOCOMOTION_TEMPLATEHELPERS_H

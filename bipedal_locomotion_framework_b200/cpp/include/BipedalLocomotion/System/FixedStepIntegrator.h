// FixedStepIntegrator.h -- kept as an include path of the reference; the class lives in Integrator.h.
#include <BipedalLocomotion/System/Integrator.h>

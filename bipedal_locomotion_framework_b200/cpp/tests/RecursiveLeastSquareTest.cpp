/**
 * @file RecursiveLeastSquareTest.cpp
 * The reference's estimator test (src/Estimators/tests/RecursiveLeastSquareTest.cpp:35-142) against
 * the GPU-backed RecursiveLeastSquare: a two-output model y = [x x^2; sin x cos x] p + noise,
 * 10 000 steps, parameters recovered within 0.1 %.  The reference loads the four parameters of
 * src/Estimators/tests/config.ini through the YARP handler; here the same values go through
 * StdImplementation.  Needs a CUDA device.
 */
#ifdef BLF_HAVE_CATCH2
#include <catch2/catch.hpp>
#else
#include "catch_shim.h"
#endif

#include <cmath>
#include <random>

#include <BipedalLocomotion/Estimators/RecursiveLeastSquare.h>
#include <BipedalLocomotion/ParametersHandler/StdImplementation.h>

using namespace BipedalLocomotion::Estimators;
using namespace BipedalLocomotion::ParametersHandler;

namespace
{
struct TwoOutputModel
{
    double p1, p2, x{0};
    std::mt19937 gen{42};
    std::normal_distribution<> noise{0, 0.5};

    iDynTree::MatrixDynSize regressor() const
    {
        iDynTree::MatrixDynSize r(2, 2);
        r(0, 0) = x;
        r(0, 1) = x * x;
        r(1, 0) = std::sin(x);
        r(1, 1) = std::cos(x);
        return r;
    }
    iDynTree::VectorDynSize output()
    {
        const iDynTree::MatrixDynSize r = regressor();
        iDynTree::VectorDynSize y(2);
        y(0) = r(0, 0) * p1 + r(0, 1) * p2 + noise(gen);
        y(1) = r(1, 0) * p1 + r(1, 1) * p2 + noise(gen);
        return y;
    }
};
} // namespace

TEST_CASE("Recursive Least Square")
{
    TwoOutputModel model{43.2, 12.2};

    auto handler = std::make_shared<StdImplementation>();
    handler->setParameter("lambda", 1.0);
    handler->setParameter("measurement_covariance", std::vector<double>{0.5, 0.5});
    handler->setParameter("state", std::vector<double>{0.0, 0.0});
    handler->setParameter("state_covariance", std::vector<double>{10.0, 10.0});

    RecursiveLeastSquare estimator;
    REQUIRE_FALSE(estimator.advance()); // neither initialised nor given a regressor
    REQUIRE(estimator.initialize(handler));
    REQUIRE_FALSE(estimator.initialize(handler)); // already initialised
    estimator.setRegressorFunction([&model]() { return model.regressor(); });

    for (int i = 0; i < 10000; i++)
    {
        model.x = std::cos(i / 10.0);
        estimator.setMeasurements(model.output());
        REQUIRE(estimator.advance());
    }

    const double admissibleError = 0.1 / 100.0;
    REQUIRE(std::abs((estimator.parametersExpectedValue()(0) - model.p1) / model.p1) < admissibleError);
    REQUIRE(std::abs((estimator.parametersExpectedValue()(1) - model.p2) / model.p2) < admissibleError);
    REQUIRE(estimator.parametersCovarianceMatrix()(0, 0) > 0);
    REQUIRE(estimator.parametersCovarianceMatrix()(0, 0) < 10.0);
}

TEST_CASE("Recursive Least Square missing parameters")
{
    auto handler = std::make_shared<StdImplementation>();
    handler->setParameter("lambda", 1.0);
    handler->setParameter("state", std::vector<double>{0.0, 0.0});
    RecursiveLeastSquare estimator;
    REQUIRE_FALSE(estimator.initialize(handler));                       // no measurement_covariance
    REQUIRE_FALSE(estimator.initialize(std::weak_ptr<IParametersHandler>())); // expired handler
}

/**
 * @file ContinuousContactModelTest.cpp
 * Checks of the GPU-backed ContinuousContactModel facade and of the batched entry point.
 *
 * The first test case holds the three properties the reference checks for this model
 * (src/ContactModels/tests/ContinousContactModelTest.cpp:32-214), with the same fixture values and
 * tolerances: Monte-Carlo surface integral vs wrench (1e-2), regressor * [k; b] vs wrench (1e-7),
 * central finite difference of the wrench vs f + g a (step 1e-6, 1e-4).  The reference writes them
 * with Eigen; Eigen is not available here, so the small vector algebra is spelled out.
 *
 * Needs a CUDA device (there is no CPU evaluation path).
 */
#ifdef BLF_HAVE_CATCH2
#include <catch2/catch.hpp>
#else
#include "catch_shim.h"
#endif

#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include <iDynTree/Core/SpatialAcc.h>

#include <BipedalLocomotion/ContactModels/ContinuousContactModel.h>
#include <BipedalLocomotion/ContactModels/ContinuousContactModelBatch.h>
#include <BipedalLocomotion/ParametersHandler/IniFile.h>
#include <BipedalLocomotion/ParametersHandler/StdImplementation.h>

using namespace iDynTree;
using namespace BipedalLocomotion::ContactModels;
using namespace BipedalLocomotion::ParametersHandler;
using BipedalLocomotion::GenericContainer::DeviceSoA;

namespace
{
constexpr double kSpring = 2000.0;
constexpr double kDamper = 100.0;
constexpr double kLength = 0.12;
constexpr double kWidth = 0.09;

std::shared_ptr<IParametersHandler> testHandler()
{
    std::shared_ptr<IParametersHandler> handler = std::make_shared<StdImplementation>();
    handler->setParameter("spring_coeff", kSpring);
    handler->setParameter("damper_coeff", kDamper);
    handler->setParameter("length", kLength);
    handler->setParameter("width", kWidth);
    return handler;
}

Transform testPose()
{
    Transform t{Transform::Identity()};
    t.setRotation(Rotation::RPY(-0.15, 0.2, 0.1));
    t.setPosition(Position(-0.02, 0.01, 0.005));
    return t;
}

Twist randomTwist(std::mt19937& gen)
{
    std::uniform_real_distribution<double> u(-1.0, 1.0); // range of Eigen's setRandom()
    Twist tw{Twist::Zero()};
    for (int i = 0; i < 6; ++i) tw(i) = u(gen);
    return tw;
}

template <typename A, typename B> void requireClose(const A& a, const B& b, double tol)
{
    REQUIRE(a.size() == b.size());
    for (std::size_t i = 0; i < a.size(); ++i) REQUIRE(std::abs(a[i] - b[i]) <= tol);
}
} // namespace

TEST_CASE("Continuous Contact")
{
    std::mt19937 gen(42);
    const Transform world_T_link = testPose();
    const Transform nullForceTransform{Transform::Identity()};
    const Twist linkVelocity = randomTwist(gen);

    ContinuousContactModel model;
    REQUIRE(model.initialize(testHandler()));
    model.setState(linkVelocity, world_T_link);
    model.setNullForceTransform(nullForceTransform);

    SECTION("Test contact wrench")
    {
        // numerical surface integral with the Monte-Carlo method
        std::default_random_engine generator;
        generator.seed(42);
        std::uniform_real_distribution<double> xAxis(-kLength / 2, kLength / 2);
        std::uniform_real_distribution<double> yAxis(-kWidth / 2, kWidth / 2);
        constexpr unsigned int samples = 10000;

        std::array<double, 6> sum{};
        for (unsigned int i = 0; i < samples; ++i)
        {
            const double x = xAxis(generator);
            const double y = yAxis(generator);
            const Force f = model.getForceAtPoint(x, y);
            const Torque t = model.getTorqueGeneratedAtPoint(x, y);
            for (int c = 0; c < 3; ++c)
            {
                sum[c] += f(c);
                sum[3 + c] += t(c);
            }
        }
        const double scale = kLength * kWidth * std::abs(world_T_link.getRotation()(2, 2)) / samples;
        Wrench numerical;
        for (int c = 0; c < 6; ++c) numerical(c) = sum[c] * scale;

        requireClose(numerical, model.getContactWrench(), 1e-2);
    }

    SECTION("Test regressor")
    {
        const MatrixDynSize regressor = model.getRegressor();
        REQUIRE(regressor.rows() == 6);
        REQUIRE(regressor.cols() == 2);
        Wrench fromRegressor;
        for (int r = 0; r < 6; ++r) fromRegressor(r) = regressor(r, 0) * kSpring + regressor(r, 1) * kDamper;
        requireClose(fromRegressor, model.getContactWrench(), 1e-7);
    }

    SECTION("Test contact dynamics")
    {
        SpatialAcc acceleration;
        for (unsigned int i = 0; i < acceleration.size(); ++i) acceleration(i) = 1;
        const double dt = 1e-6;

        // f + g a at the nominal state
        Vector6 rate = model.getAutonomousDynamics();
        const Matrix6x6& g = model.getControlMatrix();
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) rate(r) += g(r, c) * acceleration(c);

        // propagate pose (mixed representation, constant twist) and velocity by -dt and +dt
        Wrench wrenchAt[2];
        const double sign[2] = {-1.0, 1.0};
        for (int k = 0; k < 2; ++k)
        {
            Position p;
            AngularMotionVector3 rotVec;
            Twist tw;
            for (int c = 0; c < 3; ++c)
            {
                p(c) = world_T_link.getPosition()(c) + sign[k] * linkVelocity(c) * dt;
                rotVec(c) = sign[k] * linkVelocity(3 + c) * dt;
            }
            for (int c = 0; c < 6; ++c) tw(c) = linkVelocity(c) + sign[k] * acceleration(c) * dt;
            Transform t;
            t.setPosition(p);
            t.setRotation(rotVec.exp() * world_T_link.getRotation()); // R(t +- dt) = exp(+-S(w) dt) R(t)
            model.setState(tw, t);
            model.setNullForceTransform(nullForceTransform);
            wrenchAt[k] = model.getContactWrench();
        }
        Vector6 numerical;
        for (int c = 0; c < 6; ++c) numerical(c) = (wrenchAt[1](c) - wrenchAt[0](c)) / (2 * dt);
        requireClose(numerical, rate, 1e-4);
    }
}

TEST_CASE("Initialization failures and lazy cache")
{
    SECTION("Missing key, wrong type, expired handler")
    {
        const char* keys[] = {"length", "width", "spring_coeff", "damper_coeff"};
        for (const char* missing : keys)
        {
            std::shared_ptr<IParametersHandler> h = std::make_shared<StdImplementation>();
            for (const char* k : keys)
                if (std::string(k) != missing) h->setParameter(k, 0.1);
            ContinuousContactModel m;
            REQUIRE_FALSE(m.initialize(h));
        }
        std::shared_ptr<IParametersHandler> h = testHandler();
        h->setParameter("length", 1); // int where a double is required
        ContinuousContactModel m;
        REQUIRE_FALSE(m.initialize(h));
        std::weak_ptr<IParametersHandler> expired;
        REQUIRE_FALSE(m.initialize(expired));
    }

    SECTION("Defaults evaluate to zero")
    {
        // identity transforms, zero twist: every output is exactly zero whatever the parameters
        ContinuousContactModel m;
        REQUIRE(m.initialize(testHandler()));
        for (int i = 0; i < 6; ++i) REQUIRE(m.getContactWrench()(i) == 0.0);
        for (int i = 0; i < 6; ++i) REQUIRE(m.getAutonomousDynamics()(i) == 0.0);
    }

    SECTION("Coefficient written through the reference keeps the cached wrench")
    {
        std::mt19937 gen(7);
        ContinuousContactModel m;
        REQUIRE(m.initialize(testHandler()));
        m.setState(randomTwist(gen), testPose());
        const Wrench before = m.getContactWrench();
        m.springCoeff() = 2 * kSpring; // does not invalidate (reference quirk)
        for (int i = 0; i < 6; ++i) REQUIRE(m.getContactWrench()(i) == before(i));
        m.setNullForceTransform(Transform::Identity()); // any setter does
        bool changed = false;
        for (int i = 0; i < 6; ++i) changed = changed || (m.getContactWrench()(i) != before(i));
        REQUIRE(changed);
    }
}

TEST_CASE("Per-instance latency")
{
    // not a correctness check: prints what one setState + getter costs through the GPU-only facade
    std::mt19937 gen(11);
    ContinuousContactModel m;
    REQUIRE(m.initialize(testHandler()));
    const Transform pose = testPose();
    double sink = 0;
    for (int warm = 0; warm < 50; ++warm)
    {
        m.setState(randomTwist(gen), pose);
        sink += m.getContactWrench()(2);
    }
    const int iterations = 2000;
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iterations; ++i)
    {
        m.setState(randomTwist(gen), pose);
        sink += m.getContactWrench()(2);
    }
    const auto t1 = std::chrono::steady_clock::now();
    for (int i = 0; i < iterations; ++i)
    {
        m.setState(randomTwist(gen), pose);
        sink += m.getContactWrench()(2) + m.getAutonomousDynamics()(1) + m.getControlMatrix()(0, 0);
    }
    const auto t2 = std::chrono::steady_clock::now();
    const double us1 = std::chrono::duration<double, std::micro>(t1 - t0).count() / iterations;
    const double us3 = std::chrono::duration<double, std::micro>(t2 - t1).count() / iterations;
    std::printf("per-instance facade: setState + getContactWrench %.1f us; + getAutonomousDynamics + "
                "getControlMatrix %.1f us (sink %g)\n", us1, us3, sink);
    REQUIRE(us1 < 1000.0);
}

TEST_CASE("Batched entry point")
{
    std::mt19937 gen(123);
    constexpr std::size_t n = 1000;
    std::uniform_real_distribution<double> ang(-0.3, 0.3), pos(-0.05, 0.05), yaw(-3.0, 3.0);
    std::vector<Twist> twists(n);
    std::vector<Transform> poses(n), nulls(n);
    for (std::size_t i = 0; i < n; ++i)
    {
        twists[i] = randomTwist(gen);
        poses[i].setRotation(Rotation::RPY(ang(gen), ang(gen), yaw(gen)));
        poses[i].setPosition(Position(pos(gen), pos(gen), pos(gen)));
        nulls[i].setRotation(Rotation::RPY(0.1 * ang(gen), 0.1 * ang(gen), yaw(gen)));
        nulls[i].setPosition(Position(pos(gen), pos(gen), pos(gen)));
    }

    ContinuousContactModelBatch batch(0);
    REQUIRE(batch.initialize(testHandler()));

    SECTION("Arrays of objects match a loop over ContactModel instances")
    {
        std::vector<Wrench> wrenches(n);
        std::vector<Vector6> autodyn(n);
        std::vector<Matrix6x6> ctrl(n);
        REQUIRE(batch.evaluate(n, twists.data(), poses.data(), nulls.data(), nullptr,
                               ContinuousContactModelBatch::All, wrenches.data(), autodyn.data(),
                               ctrl.data()));
        ContinuousContactModel model;
        REQUIRE(model.initialize(testHandler()));
        for (std::size_t i = 0; i < n; i += 97)
        {
            model.setState(twists[i], poses[i]);
            model.setNullForceTransform(nulls[i]);
            // two different kernels (batch vs single-state): equal to the parity tolerance,
            // norm-wise per 3-vector block
            auto blockClose = [](const double* a, const double* b) {
                double num = 0, den = 0;
                for (int c = 0; c < 3; ++c)
                {
                    num = std::max(num, std::fabs(a[c] - b[c]));
                    den = std::max(den, std::fabs(b[c]));
                }
                return num <= 1e-12 * den;
            };
            REQUIRE(blockClose(wrenches[i].data(), model.getContactWrench().data()));
            REQUIRE(blockClose(wrenches[i].data() + 3, model.getContactWrench().data() + 3));
            REQUIRE(blockClose(autodyn[i].data(), model.getAutonomousDynamics().data()));
            REQUIRE(blockClose(autodyn[i].data() + 3, model.getAutonomousDynamics().data() + 3));
            for (int r = 0; r < 6; ++r)
                REQUIRE(blockClose(ctrl[i].data() + 6 * r + 3 * (r / 3),
                                   model.getControlMatrix().data() + 6 * r + 3 * (r / 3)));
        }
        // structural zeros of g are +0.0
        for (std::size_t i = 0; i < n; ++i)
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 6; ++c)
                {
                    const bool structural = (r < 3 && c != r) || (r >= 3 && c < 3);
                    if (structural) REQUIRE((ctrl[i](r, c) == 0.0 && !std::signbit(ctrl[i](r, c))));
                }
    }

    SECTION("Per-contact parameter table from the on-disk configuration format")
    {
        // four contacts, parameters from a `.ini` group (ParametersHandler/IniFile.h)
        const std::string ini = "[CONTACT_PARAMETERS]\n"
                                "length        (0.12, 0.15, 0.30, 0.08)\n"
                                "width         (0.09, 0.10, 0.15, 0.04)\n"
                                "spring_coeff  (2000.0, 50000.0, 1e6, 1e3)\n"
                                "damper_coeff  (100.0, 300.0, 1e4, 10.0)\n";
        auto file = std::make_shared<StdImplementation>();
        REQUIRE(loadIniString(ini, *file));
        DeviceSoA table;
        REQUIRE(batch.loadParameterTable(file->getGroup("CONTACT_PARAMETERS"), table));
        REQUIRE(table.planes() == 4);
        REQUIRE(table.size() == 4);
        REQUIRE_FALSE(batch.loadParameterTable(file->getGroup("MISSING"), table)); // expired group
        auto shortTable = std::make_shared<StdImplementation>();
        REQUIRE(loadIniString("length (0.1, 0.2)\nwidth (0.1)\nspring_coeff (1.0, 2.0)\n"
                              "damper_coeff (1.0, 2.0)\n", *shortTable));
        REQUIRE_FALSE(batch.loadParameterTable(shortTable, table));                // ragged columns
        REQUIRE(table.size() == 4);                                                 // untouched

        const std::size_t m = 4;
        auto dev = batch.device();
        DeviceSoA states(dev, ContinuousContactModelBatch::NumberOfPlanes, m), wrenchPlanes(dev, 6, m);
        REQUIRE(states.uploadRows(ContinuousContactModelBatch::LinearVelocity, 6,
                                  reinterpret_cast<const double*>(twists.data())));
        REQUIRE(states.uploadRows(ContinuousContactModelBatch::Position, 12,
                                  reinterpret_cast<const double*>(poses.data())));
        REQUIRE(states.uploadRows(ContinuousContactModelBatch::NullForcePosition, 12,
                                  reinterpret_cast<const double*>(nulls.data())));
        REQUIRE(batch.evaluate(states, &table, ContinuousContactModelBatch::ContactWrench, &wrenchPlanes,
                               nullptr, nullptr, nullptr));
        std::vector<Wrench> got(m);
        REQUIRE(wrenchPlanes.downloadRows(0, 6, reinterpret_cast<double*>(got.data())));
        const double L[4] = {0.12, 0.15, 0.30, 0.08}, W[4] = {0.09, 0.10, 0.15, 0.04};
        const double K[4] = {2000.0, 50000.0, 1e6, 1e3}, B[4] = {100.0, 300.0, 1e4, 10.0};
        for (std::size_t i = 0; i < m; ++i)
        {
            auto h = std::make_shared<StdImplementation>();
            h->setParameter("length", L[i]);
            h->setParameter("width", W[i]);
            h->setParameter("spring_coeff", K[i]);
            h->setParameter("damper_coeff", B[i]);
            ContinuousContactModel model;
            REQUIRE(model.initialize(h));
            model.setState(twists[i], poses[i]);
            model.setNullForceTransform(nulls[i]);
            for (int c = 0; c < 6; ++c)
            {
                const double ref = model.getContactWrench()(c);
                REQUIRE(std::fabs(got[i](c) - ref) <= 1e-12 * std::max(std::fabs(ref), 1e-300));
            }
        }
    }

    SECTION("Device SoA container and the rollout arg-min")
    {
        auto dev = batch.device();
        DeviceSoA states(dev, ContinuousContactModelBatch::NumberOfPlanes, n);
        REQUIRE(states.valid());
        REQUIRE(states.uploadRows(ContinuousContactModelBatch::LinearVelocity, 6,
                                  reinterpret_cast<const double*>(twists.data())));
        REQUIRE(states.uploadRows(ContinuousContactModelBatch::Position, 12,
                                  reinterpret_cast<const double*>(poses.data())));
        REQUIRE(states.uploadRows(ContinuousContactModelBatch::NullForcePosition, 12,
                                  reinterpret_cast<const double*>(nulls.data())));
        DeviceSoA wrenchPlanes(dev, 6, n);
        REQUIRE(batch.evaluate(states, nullptr, ContinuousContactModelBatch::ContactWrench, &wrenchPlanes,
                               nullptr, nullptr, nullptr));
        std::vector<double> w(n * 6);
        REQUIRE(wrenchPlanes.downloadRows(0, 6, w.data()));

        std::vector<Wrench> wrenches(n);
        REQUIRE(batch.evaluate(n, twists.data(), poses.data(), nulls.data(), nullptr,
                               ContinuousContactModelBatch::ContactWrench, wrenches.data(), nullptr, nullptr));
        for (std::size_t i = 0; i < n; ++i)
            for (int c = 0; c < 6; ++c) REQUIRE(std::abs(w[i * 6 + c] - wrenches[i](c)) <= 1e-12 * (1 + std::abs(w[i * 6 + c])));

        // 10 rollouts of 100 evaluations: cost and arg-min against a host loop over the wrenches
        Wrench reference;
        reference(2) = 30.0;
        ContinuousContactModelBatch::RolloutResult best{};
        REQUIRE(batch.rolloutCostArgmin(states, nullptr, 100, reference, 1.0, 10.0, best));
        double bestCost = 1e300;
        std::int64_t bestIndex = -1;
        for (std::size_t r = 0; r < n / 100; ++r)
        {
            double cost = 0;
            for (std::size_t e = 0; e < 100; ++e)
            {
                const double* x = &w[(r * 100 + e) * 6];
                double qf = 0, qt = 0;
                for (int c = 0; c < 3; ++c)
                {
                    qf += (x[c] - reference(c)) * (x[c] - reference(c));
                    qt += (x[3 + c] - reference(3 + c)) * (x[3 + c] - reference(3 + c));
                }
                cost += qf + 10.0 * qt;
            }
            if (cost < bestCost)
            {
                bestCost = cost;
                bestIndex = static_cast<std::int64_t>(r);
            }
        }
        REQUIRE(best.index == bestIndex);
        REQUIRE(std::abs(best.cost - bestCost) <= 1e-10 * bestCost);
    }
}

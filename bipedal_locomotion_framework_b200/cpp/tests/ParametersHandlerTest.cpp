/**
 * @file ParametersHandlerTest.cpp
 * Behaviour of StdImplementation that the contact-model path relies on; the cases are those of the
 * reference's src/ParametersHandler/tests/ParametersHandlerTest.cpp:25-117 plus the strict-typing
 * and missing-key failures ContinuousContactModel::initialize depends on.  Pure host code.
 */
#ifdef BLF_HAVE_CATCH2
#include <catch2/catch.hpp>
#else
#include "catch_shim.h"
#endif

#include <any>
#include <array>
#include <iostream>
#include <unordered_map>

#include <iDynTree/Core/VectorDynSize.h>

#include <BipedalLocomotion/ParametersHandler/IniFile.h>
#include <BipedalLocomotion/ParametersHandler/StdImplementation.h>

using namespace BipedalLocomotion::ParametersHandler;

TEST_CASE("Get parameters")
{
    const std::vector<int> fibonacciNumbers{1, 1, 2, 3, 5, 8, 13, 21};
    const std::vector<std::string> donaldsNephews{"Huey", "Dewey", "Louie"};
    std::shared_ptr<StdImplementation> originalHandler = std::make_shared<StdImplementation>();
    IParametersHandler::shared_ptr parameterHandler = originalHandler;

    parameterHandler->setParameter("answer_to_the_ultimate_question_of_life", 42);
    parameterHandler->setParameter("pi", 3.14);
    parameterHandler->setParameter("Fibonacci Numbers", std::vector<int>{1, 1, 2, 3, 5, 8, 13, 21});
    parameterHandler->setParameter("John", "Smith");

    SECTION("Get integer")
    {
        int element;
        REQUIRE(parameterHandler->getParameter("answer_to_the_ultimate_question_of_life", element));
        REQUIRE(element == 42);
    }

    SECTION("Get Double")
    {
        double element;
        REQUIRE(parameterHandler->getParameter("pi", element));
        REQUIRE(element == 3.14);
    }

    SECTION("Get String")
    {
        std::string element;
        REQUIRE(parameterHandler->getParameter("John", element));
        REQUIRE(element == "Smith");
    }

    SECTION("Get Vector")
    {
        std::vector<int> element;
        // the reference's contract (IParametersHandler.h:129-139): Fixed (default) needs a destination of
        // the right size already, Resizable resizes it
        REQUIRE_FALSE(parameterHandler->getParameter("Fibonacci Numbers", element));
        REQUIRE(parameterHandler->getParameter("Fibonacci Numbers", element, BipedalLocomotion::GenericContainer::VectorResizeMode::Resizable));
        REQUIRE(element == std::vector<int>{1, 1, 2, 3, 5, 8, 13, 21});
        std::vector<int> sized(8, 0);
        REQUIRE(parameterHandler->getParameter("Fibonacci Numbers", sized));
        REQUIRE(sized == element);
    }

    SECTION("Get Vector into any container a GenericContainer::Vector can view")
    {
        using BipedalLocomotion::GenericContainer::VectorResizeMode;
        namespace GC = BipedalLocomotion::GenericContainer;
        parameterHandler->setParameter("gains", std::vector<double>{1.5, -2.0, 0.25});
        // fixed-size destinations: std::array and a plain array of the right size succeed, a wrong size fails
        std::array<double, 3> arr{};
        REQUIRE(parameterHandler->getParameter("gains", arr));
        REQUIRE((arr[0] == 1.5 && arr[1] == -2.0 && arr[2] == 0.25));
        std::array<double, 4> wrong{};
        REQUIRE_FALSE(parameterHandler->getParameter("gains", wrong));
        REQUIRE_FALSE(parameterHandler->getParameter("gains", wrong, VectorResizeMode::Resizable)); // cannot resize
        double plain[3] = {0, 0, 0};
        REQUIRE(parameterHandler->getParameter("gains", plain));
        REQUIRE(plain[2] == 0.25);
        // iDynTree::VectorDynSize: resized on request, as reference user code expects
        iDynTree::VectorDynSize dyn;
        REQUIRE_FALSE(parameterHandler->getParameter("gains", dyn));
        REQUIRE(parameterHandler->getParameter("gains", dyn, VectorResizeMode::Resizable));
        REQUIRE((dyn.size() == 3 && dyn(1) == -2.0));
        // the view itself, and strict typing through it
        std::vector<double> owner(3);
        auto view = GC::make_vector(owner);
        REQUIRE(parameterHandler->getParameter("gains", view));
        REQUIRE(owner[0] == 1.5);
        std::array<int, 3> ints{};
        REQUIRE_FALSE(parameterHandler->getParameter("gains", ints));          // stored as double
        std::array<int, 8> fib{};
        REQUIRE(parameterHandler->getParameter("Fibonacci Numbers", fib));
        REQUIRE(fib[7] == 21);
        // set from a non-std::vector container
        parameterHandler->setParameter("from_array", arr);
        std::vector<double> back;
        REQUIRE(parameterHandler->getParameter("from_array", back, VectorResizeMode::Resizable));
        REQUIRE(back == std::vector<double>{1.5, -2.0, 0.25});
        static_assert(GC::is_vector<GC::Vector<double>>::value && !GC::is_vector<std::vector<double>>::value, "traits");
        static_assert(GC::is_vector_constructible<std::array<double, 3>>::value
                          && !GC::is_vector_constructible<std::string>::value && !GC::is_vector_constructible<double>::value,
                      "traits");
    }

    SECTION("Strict typing and missing keys")
    {
        double asDouble;
        int asInt;
        REQUIRE_FALSE(parameterHandler->getParameter("answer_to_the_ultimate_question_of_life", asDouble));
        REQUIRE_FALSE(parameterHandler->getParameter("pi", asInt));
        REQUIRE_FALSE(parameterHandler->getParameter("not there", asDouble));
    }

    SECTION("Set/Get Group")
    {
        IParametersHandler::shared_ptr newGroup = std::make_shared<StdImplementation>();
        REQUIRE(parameterHandler->setGroup("CARTOONS", newGroup));
        IParametersHandler::shared_ptr groupHandler = parameterHandler->getGroup("CARTOONS").lock();
        REQUIRE(groupHandler);
        groupHandler->setParameter("Donald's nephews", std::vector<std::string>{"Huey", "Dewey", "Louie"});
        std::vector<std::string> element;
        REQUIRE(groupHandler->getParameter("Donald's nephews", element, BipedalLocomotion::GenericContainer::VectorResizeMode::Resizable));
        REQUIRE(element == std::vector<std::string>{"Huey", "Dewey", "Louie"});
        REQUIRE_FALSE(parameterHandler->getGroup("NO SUCH GROUP").lock());
    }

    SECTION("is Empty")
    {
        IParametersHandler::shared_ptr newGroup = std::make_shared<StdImplementation>();
        REQUIRE(parameterHandler->setGroup("CARTOONS", newGroup));
        IParametersHandler::shared_ptr groupHandler = parameterHandler->getGroup("CARTOONS").lock();
        REQUIRE(groupHandler);
        REQUIRE(groupHandler->isEmpty());
        groupHandler->setParameter("Donald's nephews", std::vector<std::string>{"Huey", "Dewey", "Louie"});
        REQUIRE_FALSE(groupHandler->isEmpty());
    }

    SECTION("Print content")
    {
        std::cout << "Parameters: " << parameterHandler->toString() << std::endl;
    }

    SECTION("Set from object")
    {
        std::unordered_map<std::string, std::any> object;
        object["value"] = std::make_any<int>(10);
        originalHandler->set(object);
        int expected;
        REQUIRE(parameterHandler->getParameter("value", expected));
        REQUIRE(expected == 10);
    }

    SECTION("Clear")
    {
        REQUIRE_FALSE(parameterHandler->isEmpty());
        parameterHandler->clear();
        REQUIRE(parameterHandler->isEmpty());
    }

    SECTION("Set from ini text")
    {
        // What the reference's "Set from RF" section checks
        // (src/ParametersHandler/tests/ParametersHandlerYarpTest.cpp:133-190) on the content of its
        // fixture src/ParametersHandler/tests/config.ini, read here without YARP.
        const std::string ini = "answer_to_the_ultimate_question_of_life 42\n"
                                "pi                                      3.14\n"
                                "John                                    Smith\n"
                                "\"Fibonacci Numbers\"                     (1, 1, 2, 3, 5, 8, 13, 21)\n"
                                "\n"
                                "[CARTOONS]\n"
                                "\"Donald's nephews\"                      (\"Huey\", \"Dewey\", \"Louie\")\n"
                                "Fibonacci_Numbers                       (1, 1, 2, 3, 5, 8, 13, 21)\n"
                                "John                                    Doe\n";
        parameterHandler->clear();
        REQUIRE(parameterHandler->isEmpty());
        REQUIRE(loadIniString(ini, *parameterHandler));
        {
            int element;
            REQUIRE(parameterHandler->getParameter("answer_to_the_ultimate_question_of_life", element));
            REQUIRE(element == 42);
            double wrongType;
            REQUIRE_FALSE(parameterHandler->getParameter("answer_to_the_ultimate_question_of_life", wrongType));
        }
        {
            double element;
            REQUIRE(parameterHandler->getParameter("pi", element));
            REQUIRE(element == 3.14);
        }
        {
            std::string element;
            REQUIRE(parameterHandler->getParameter("John", element));
            REQUIRE(element == "Smith");
        }
        {
            std::vector<int> element;
            REQUIRE(parameterHandler->getParameter("Fibonacci Numbers", element, BipedalLocomotion::GenericContainer::VectorResizeMode::Resizable));
            REQUIRE(element == fibonacciNumbers);
        }
        IParametersHandler::shared_ptr cartoonsGroup = parameterHandler->getGroup("CARTOONS").lock();
        REQUIRE(cartoonsGroup);
        {
            std::vector<std::string> element;
            REQUIRE(cartoonsGroup->getParameter("Donald's nephews", element, BipedalLocomotion::GenericContainer::VectorResizeMode::Resizable));
            REQUIRE(element == donaldsNephews);
        }
        {
            std::vector<int> element;
            REQUIRE(cartoonsGroup->getParameter("Fibonacci_Numbers", element, BipedalLocomotion::GenericContainer::VectorResizeMode::Resizable));
            REQUIRE(element == fibonacciNumbers);
        }
        {
            std::string element;
            REQUIRE(cartoonsGroup->getParameter("John", element));
            REQUIRE(element == "Doe");
        }
    }

    SECTION("Ini grammar: comments, bare lists, booleans, exponents, errors")
    {
        StdImplementation h;
        REQUIRE(loadIniString("# a comment\n// another\n  rho 1e-2   # trailing\n"
                              "gains 1.5 2 2.5\nflags (true, false)\nuse_cuda true\nempty ()\n"
                              "name \"two words\"\n[CONTACT_PARAMETERS]\nlength (0.12, 0.15)\n",
                              h));
        double rho;
        REQUIRE(h.getParameter("rho", rho));
        REQUIRE(rho == 0.01);
        std::vector<double> gains;
        REQUIRE(h.getParameter("gains", gains, BipedalLocomotion::GenericContainer::VectorResizeMode::Resizable)); // mixed int/double list -> doubles
        REQUIRE((gains == std::vector<double>{1.5, 2.0, 2.5}));
        std::vector<bool> flags;
        REQUIRE(h.getParameter("flags", flags));
        REQUIRE((flags == std::vector<bool>{true, false}));
        bool useCuda = false;
        REQUIRE(h.getParameter("use_cuda", useCuda));
        REQUIRE(useCuda);
        std::vector<int> empty{1};
        REQUIRE(h.getParameter("empty", empty, BipedalLocomotion::GenericContainer::VectorResizeMode::Resizable));
        REQUIRE(empty.empty());
        std::string name;
        REQUIRE(h.getParameter("name", name));
        REQUIRE(name == "two words");
        auto table = h.getGroup("CONTACT_PARAMETERS").lock();
        REQUIRE(table);
        std::vector<double> length;
        REQUIRE(table->getParameter("length", length, BipedalLocomotion::GenericContainer::VectorResizeMode::Resizable));
        REQUIRE((length == std::vector<double>{0.12, 0.15}));

        StdImplementation bad;
        REQUIRE_FALSE(loadIniString("key \"unterminated\n", bad));
        REQUIRE_FALSE(loadIniString("key (1, 2\n", bad));
        REQUIRE_FALSE(loadIniString("lonely_key\n", bad));
        REQUIRE_FALSE(loadIniString("[]\n", bad));
        REQUIRE_FALSE(loadIniFile("/nonexistent/config.ini", bad));
    }
}

/**
 * @file ContactWrenchTest.cpp
 * Host-only checks of System::ContactWrench (the reference has no test for it; behaviour from
 * src/System/src/ContactWrench.cpp:13-35): shared ownership, weak hand-out, mutable frame index.
 */
#ifdef BLF_HAVE_CATCH2
#include <catch2/catch.hpp>
#else
#include "catch_shim.h"
#endif

#include <vector>

#include <BipedalLocomotion/ContactModels/ContinuousContactModel.h>
#include <BipedalLocomotion/System/ContactWrench.h>

using namespace BipedalLocomotion::ContactModels;
using namespace BipedalLocomotion::System;

TEST_CASE("ContactWrench holds the frame index and shares the model")
{
    auto model = std::make_shared<ContinuousContactModel>(); // no device is touched before initialize()
    std::weak_ptr<ContactModel> observer;
    {
        std::vector<ContactWrench> contacts;
        contacts.emplace_back(iDynTree::FrameIndex(7), model);
        contacts.emplace_back(iDynTree::FrameIndex(11), model);
        REQUIRE(contacts[0].index() == 7);
        contacts[0].index() = 9;
        const ContactWrench& first = contacts[0];
        REQUIRE(first.index() == 9);
        REQUIRE(model.use_count() == 3);

        observer = first.contactModel();
        auto locked = observer.lock();
        REQUIRE(locked);
        REQUIRE(locked.get() == model.get());
        locked.reset();

        model.reset(); // the holders keep it alive
        REQUIRE(observer.lock());
    }
    REQUIRE_FALSE(observer.lock()); // the last holder is gone
}

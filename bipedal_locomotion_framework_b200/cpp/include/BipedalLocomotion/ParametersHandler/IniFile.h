/**
 * @file IniFile.h
 * Reader for the reference's on-disk configuration format: the YARP `.ini` files its handlers are
 * filled from (src/ParametersHandler/YarpImplementation/src/YarpImplementation.cpp:115-144 via
 * yarp::os::ResourceFinder; fixture src/ParametersHandler/tests/config.ini:1-9).  YARP itself is a
 * third-party dependency that is absent here, so the subset of its text format those files use is
 * parsed directly into any IParametersHandler:
 *
 *     # comment            // comment
 *     key value                      scalar: 42 -> int, 3.14 / 1e3 -> double, true/false -> bool,
 *     "quoted key" "quoted value"            anything else -> string
 *     key (1, 2, 3)                  list -> std::vector<int|double|bool|std::string>; commas and
 *     key 1 2 3                      blanks both separate; several bare values also form a list
 *     [GROUP]                        every following line goes to the group GROUP
 *
 * Typing is as strict as the reference's (YarpUtilities/Helper.tpp:30-38): an int cannot be read as
 * a double.  One relaxation: a numeric list that mixes ints and doubles is stored as
 * std::vector<double> (YARP would reject it under either type).
 *
 * Per-contact parameter tables: a group holding the four keys of
 * ContinuousContactModel::initialize as equally long lists, e.g.
 *
 *     [CONTACT_PARAMETERS]
 *     length        (0.12, 0.15)
 *     width         (0.09, 0.10)
 *     spring_coeff  (2000.0, 50000.0)
 *     damper_coeff  (100.0, 300.0)
 *
 * is turned into the four device planes the batched evaluation takes by
 * ContinuousContactModelBatch::loadParameterTable.
 */
#ifndef BIPEDAL_LOCOMOTION_PARAMETERS_HANDLER_INI_FILE_H
#define BIPEDAL_LOCOMOTION_PARAMETERS_HANDLER_INI_FILE_H

#include <memory>
#include <string>

#include <BipedalLocomotion/ParametersHandler/IParametersHandler.h>

namespace BipedalLocomotion
{
namespace ParametersHandler
{

/** Parse `text` into `handler` (groups become StdImplementation handlers).  false + std::cerr
 * message with the line number on a malformed line; what was parsed before stays set. */
bool loadIniString(const std::string& text, IParametersHandler& handler);

/** Same for a file; false if it cannot be opened. */
bool loadIniFile(const std::string& path, IParametersHandler& handler);

} // namespace ParametersHandler
} // namespace BipedalLocomotion

#endif // BIPEDAL_LOCOMOTION_PARAMETERS_HANDLER_INI_FILE_H

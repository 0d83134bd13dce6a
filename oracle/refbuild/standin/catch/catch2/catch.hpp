// catch_shim.h -- the few Catch2 v2 macros the reference's tests use (TEST_CASE, SECTION, REQUIRE,
// REQUIRE_FALSE), for environments where Catch2 is not installed.  Semantics kept: the TEST_CASE
// body is re-run from the top once per SECTION, so every section starts from a fresh setup.
// CMake uses the real <catch2/catch.hpp> when find_package(Catch2) succeeds.
#ifndef BLF_CATCH_SHIM_H
#define BLF_CATCH_SHIM_H

#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

namespace catch_shim
{
struct TestCase
{
    const char* name;
    void (*fn)();
};
inline std::vector<TestCase>& registry()
{
    static std::vector<TestCase> r;
    return r;
}
struct Registrar
{
    Registrar(const char* name, void (*fn)()) { registry().push_back({name, fn}); }
};
struct Failure : std::exception
{
};
struct State
{
    int target = 0;   // index of the section to run in this pass
    int seen = 0;     // sections encountered so far in this pass
    int assertions = 0;
    int failures = 0;
    std::string section;
};
inline State& state()
{
    static State s;
    return s;
}
inline bool enterSection(const char* name)
{
    State& s = state();
    const bool run = (s.seen == s.target);
    if (run) s.section = name;
    ++s.seen;
    return run;
}
inline void check(bool ok, const char* expr, const char* file, int line)
{
    State& s = state();
    ++s.assertions;
    if (ok) return;
    ++s.failures;
    std::fprintf(stderr, "%s:%d: FAILED: REQUIRE( %s )  [section: %s]\n", file, line, expr,
                 s.section.c_str());
    throw Failure();
}
inline int runAll(int argc, char** argv)
{
    const char* filter = argc > 1 ? argv[1] : nullptr;
    int failedCases = 0, ran = 0;
    for (const TestCase& tc : registry())
    {
        if (filter && std::strstr(tc.name, filter) == nullptr) continue;
        ++ran;
        State& s = state();
        const int failuresBefore = s.failures;
        s.target = 0;
        for (;;)
        {
            s.seen = 0;
            s.section = "<none>";
            try
            {
                tc.fn();
            } catch (const Failure&)
            {
            } catch (const std::exception& e)
            {
                ++s.failures;
                std::fprintf(stderr, "unexpected exception in '%s': %s\n", tc.name, e.what());
            }
            if (s.target + 1 >= s.seen) break;
            ++s.target;
        }
        const bool ok = s.failures == failuresBefore;
        std::printf("%s  %s\n", ok ? "[ OK ]" : "[FAIL]", tc.name);
        if (!ok) ++failedCases;
    }
    std::printf("%d test case(s), %d assertion(s), %d failure(s)\n", ran, state().assertions,
                state().failures);
    return (failedCases == 0 && ran > 0) ? 0 : 1;
}
} // namespace catch_shim

#define BLF_CS_CAT2(a, b) a##b
#define BLF_CS_CAT(a, b) BLF_CS_CAT2(a, b)
#define TEST_CASE(name)                                                                   \
    static void BLF_CS_CAT(blf_test_fn_, __LINE__)();                                     \
    static catch_shim::Registrar BLF_CS_CAT(blf_test_reg_, __LINE__)(name, &BLF_CS_CAT(blf_test_fn_, __LINE__)); \
    static void BLF_CS_CAT(blf_test_fn_, __LINE__)()
#define SECTION(name) if (catch_shim::enterSection(name))
#define REQUIRE(...) catch_shim::check(static_cast<bool>(__VA_ARGS__), #__VA_ARGS__, __FILE__, __LINE__)
#define REQUIRE_FALSE(...) \
    catch_shim::check(!static_cast<bool>(__VA_ARGS__), "!(" #__VA_ARGS__ ")", __FILE__, __LINE__)

#ifdef CATCH_CONFIG_MAIN
int main(int argc, char** argv) { return catch_shim::runAll(argc, argv); }
#endif

#endif // BLF_CATCH_SHIM_H

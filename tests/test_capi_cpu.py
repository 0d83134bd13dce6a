"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/blf_ccm.h declares, fails loudly without a device, and the host-side handler logic
matches the reference's (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from bipedal_locomotion_framework_b200 import _capi, build
from bipedal_locomotion_framework_b200.contact_models import (ContinuousContactModel,
                                                               StdImplementation)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_cuda()
    return _capi.lib()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "blf_ccm.h")).read()
    declared = set(re.findall(r"BLF_CCM_API\s+[\w\s\*]+?\b(blf_(?:ccm|rls|sys)_\w+)\s*\(", header))
    assert declared, "no declarations parsed from include/blf_ccm.h"
    assert declared == set(_capi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/blf_ccm.h but not exported"


def test_version_and_sass_target(lib):
    assert b"sm_100a" in lib.blf_ccm_version()


def test_invalid_handle_is_rejected(lib):
    assert lib.blf_ccm_set_uniform_params(None, 0.1, 0.1, 1.0, 1.0) == _capi.ERR_INVALID_HANDLE
    assert lib.blf_ccm_launch_count(None) == -1
    assert b"invalid handle" in lib.blf_ccm_last_error()
    # the entry points of the System rows check the handle before anything else (no CUDA call is made)
    assert lib.blf_sys_mass_matrix_solve(None, 1, 6, None, None, None, None, None, None) == _capi.ERR_INVALID_HANDLE
    assert lib.blf_sys_floating_base_acceleration(None, 1, 1, 6, None, None, None, None, None, None, None, None,
                                                  None, None) == _capi.ERR_INVALID_HANDLE
    assert lib.blf_ccm_generalized_force_soa(None, 1, 1, 6, None, None, None, None, None, None,
                                             None) == _capi.ERR_INVALID_HANDLE


def test_create_fails_loudly_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = lib.blf_ccm_create(0, C.byref(h))
    assert rc == _capi.ERR_NO_DEVICE and not h
    assert b"no CPU path" in lib.blf_ccm_last_error()
    # the facade reports it the reference's way: false + message, never a silent CPU result
    handler = StdImplementation()
    for k, v in (("length", 0.12), ("width", 0.09), ("spring_coeff", 2000.0),
                 ("damper_coeff", 100.0)):
        handler.setParameter(k, v)
    m = ContinuousContactModel()
    assert m.initialize(handler) is False
    with pytest.raises(RuntimeError):
        m.getContactWrench()


# --- ParametersHandler semantics (src/ParametersHandler/tests/ParametersHandlerTest.cpp:25-117) ---

def test_std_implementation_semantics():
    h = StdImplementation()
    h.setParameter("answer_to_the_ultimate_question_of_life", 42)
    h.setParameter("pi", 3.14)
    h.setParameter("John", "Smith")
    assert h.getParameter("answer_to_the_ultimate_question_of_life", int) == (True, 42)
    assert h.getParameter("pi", float) == (True, 3.14)
    assert h.getParameter("John", str) == (True, "Smith")
    assert h.getParameter("pi", int)[0] is False          # strict any_cast typing
    assert h.getParameter("missing", float)[0] is False
    g = StdImplementation()
    assert h.setGroup("CARTOONS", g)
    assert h.getGroup("CARTOONS").isEmpty()
    g.setParameter("nephews", "Huey")
    assert not h.getGroup("CARTOONS").isEmpty()
    assert h.getGroup("nope").isEmpty()                   # missing group -> fresh empty handler
    h.set({"value": 10})
    assert h.getParameter("value", int) == (True, 10)
    assert not h.isEmpty()
    h.clear()
    assert h.isEmpty()


def test_initialize_rejects_missing_or_mistyped_keys():
    ok = {"length": 0.12, "width": 0.09, "spring_coeff": 2000.0, "damper_coeff": 100.0}
    for drop in ok:
        h = StdImplementation()
        for k, v in ok.items():
            if k != drop:
                h.setParameter(k, v)
        assert ContinuousContactModel().initialize(h) is False
    h = StdImplementation()
    for k, v in ok.items():
        h.setParameter(k, v)
    h.setParameter("length", 1)                            # int under a double key
    assert ContinuousContactModel().initialize(h) is False
    assert ContinuousContactModel().initialize(None) is False


# --- on-disk configuration format ------------------------------------

REFERENCE_CONFIG_INI = '''answer_to_the_ultimate_question_of_life 42
pi                                      3.14
John                                    Smith
"Fibonacci Numbers"                     (1, 1, 2, 3, 5, 8, 13, 21)

[CARTOONS]
"Donald's nephews"                      ("Huey", "Dewey", "Louie")
Fibonacci_Numbers                       (1, 1, 2, 3, 5, 8, 13, 21)
John                                    Doe
'''


def test_ini_reader_on_the_reference_fixture_content():
    """What ParametersHandlerYarpTest.cpp:133-190 ("Set from RF") checks on the content of
    src/ParametersHandler/tests/config.ini, read without YARP."""
    from bipedal_locomotion_framework_b200.ini import load_ini_string
    h = load_ini_string(REFERENCE_CONFIG_INI)
    assert h.getParameter("answer_to_the_ultimate_question_of_life", int) == (True, 42)
    assert h.getParameter("answer_to_the_ultimate_question_of_life", float)[0] is False   # strict
    assert h.getParameter("pi", float) == (True, 3.14)
    assert h.getParameter("John", str) == (True, "Smith")
    assert h.getParameter("Fibonacci Numbers", list) == (True, [1, 1, 2, 3, 5, 8, 13, 21])
    g = h.getGroup("CARTOONS")
    assert g.getParameter("Donald's nephews", list) == (True, ["Huey", "Dewey", "Louie"])
    assert g.getParameter("Fibonacci_Numbers", list) == (True, [1, 1, 2, 3, 5, 8, 13, 21])
    assert g.getParameter("John", str) == (True, "Doe")


def test_ini_grammar_and_parameter_table():
    import numpy as np
    from bipedal_locomotion_framework_b200.ini import load_ini_string, parameter_table
    h = load_ini_string('# c\n// c\n rho 1e-2 # t\ngains 1.5 2 2.5\nflags (true, false)\nempty ()\n'
                        'name "two words"\n[CONTACT_PARAMETERS]\nlength (0.12, 0.15)\n'
                        'width (0.09, 0.1)\nspring_coeff (2000.0, 5e4)\ndamper_coeff (100.0, 300.0)\n')
    assert h.getParameter("rho", float) == (True, 0.01)
    assert h.getParameter("gains", list) == (True, [1.5, 2.0, 2.5])
    assert h.getParameter("flags", list) == (True, [True, False])
    assert h.getParameter("empty", list) == (True, [])
    assert h.getParameter("name", str) == (True, "two words")
    t = parameter_table(h.getGroup("CONTACT_PARAMETERS"))
    assert np.array_equal(t, [[0.12, 0.15], [0.09, 0.1], [2000.0, 5e4], [100.0, 300.0]])
    for bad in ('key "unterminated\n', "key (1, 2\n", "lonely\n", "[]\n", "k ((1) 2)\n"):
        assert load_ini_string(bad) is None
    ragged = load_ini_string("length (0.1, 0.2)\nwidth (0.1)\nspring_coeff (1.0, 2.0)\n"
                             "damper_coeff (1.0, 2.0)\n")
    assert parameter_table(ragged) is None
    ints = load_ini_string("length (1, 2)\nwidth (1.0, 2.0)\nspring_coeff (1.0, 2.0)\n"
                           "damper_coeff (1.0, 2.0)\n")
    assert parameter_table(ints) is None                  # ints are not doubles (strict typing)
    assert parameter_table(None) is None


def test_header_is_plain_c99_with_no_cuda_or_torch_types(tmp_path):
    """The drop-in boundary is a C ABI: include/blf_ccm.h must compile as ISO C99 on its own (plain
    pointers and sizes in every signature -- no C++, CUDA or torch type)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "hdr.c"
    src.write_text('#include "blf_ccm.h"\nint main(void) { return blf_ccm_version() != 0 ? 0 : 1; }\n')
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "include"), str(src)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    header = open(os.path.join(ROOT, "include", "blf_ccm.h")).read()
    code = re.sub(r"/\*.*?\*/", "", header, flags=re.S)          # declarations only, comments stripped
    code = re.sub(r"//[^\n]*", "", code)
    for forbidden in ("cudaStream_t", "torch", "at::", "std::", "#include <cuda", "class ", "template"):
        assert forbidden not in code, f"{forbidden} leaks into the C ABI"


def test_every_entry_point_cites_the_reference_or_says_it_is_new():
    """Each declaration's comment block names the reference function it replaces (file:line under
    src/...) or states that it has no reference counterpart."""
    header = open(os.path.join(ROOT, "include", "blf_ccm.h")).read()
    cites = set(re.findall(r"\b\w+\.(?:cpp|h|tpp):\d+", header))
    assert len(cites) >= 12, f"only {len(cites)} distinct file:line citations in include/blf_ccm.h"
    for needed in ("ContinuousContactModel.cpp", "ContactModel.cpp", "RecursiveLeastSquare.cpp",
                   "FloatingBaseSystemKinematics.cpp", "FloatingBaseSystemDynamics.cpp"):
        assert any(c.startswith(needed) for c in cites), f"no citation of {needed}"

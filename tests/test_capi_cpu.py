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

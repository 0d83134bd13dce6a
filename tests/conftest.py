import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_unusable():
    """None when the sm_100a backend can open cuda:0, else the reason."""
    try:
        import torch
        if not torch.cuda.is_available():
            return "no CUDA device"
        import ctypes as C

        from bipedal_locomotion_framework_b200 import _capi
        h = C.c_void_p()
        rc = _capi.lib().blf_ccm_create(0, C.byref(h))
        if rc != 0:
            return "blf_ccm_create failed: " + _capi.lib().blf_ccm_last_error().decode()
        _capi.lib().blf_ccm_destroy(h)
        return None
    except Exception as e:   # missing library, driver trouble
        return f"{type(e).__name__}: {e}"


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a CPU-only box: gpu-marked tests are skipped, not failed.  On a GPU
    box they run -- and there a missing library is an error, not a skip (the product has no CPU
    path; the driver's `-m gpu` run must load libblf_ccm.so)."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    try:
        import torch
        have_cuda = torch.cuda.is_available()
    except Exception:
        have_cuda = False
    if have_cuda:
        return
    skip = pytest.mark.skip(reason="gpu test: " + (_gpu_unusable() or "no CUDA device"))
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ccm_exact_golden.npz")))


@pytest.fixture(scope="session")
def oracle():
    from oracle import ccm_oracle
    ccm_oracle.build()
    return ccm_oracle

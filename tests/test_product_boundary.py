"""The product never routes through the oracle (or any CPU evaluation): static scan of the product
tree, and the package imported with `oracle` made unimportable."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "bipedal_locomotion_framework_b200")


def _product_files():
    for base, dirs, files in os.walk(PKG):
        dirs[:] = [d for d in dirs if d not in ("lib", "__pycache__", ".pytest_cache")]
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".cpp", ".h")):
                yield os.path.join(base, f)


def test_product_sources_never_import_link_or_execute_the_oracle():
    offenders = []
    pat = re.compile(r"^\s*(?:from\s+oracle\b|import\s+oracle\b|#\s*include\s*[<\"][^>\"]*oracle)|"
                     r"libccm_oracle|libblf_reference|ref_binding|ccm_oracle\.|sys_oracle\.|rls_oracle\.", re.M)
    for path in _product_files():
        if os.sep + os.path.join("cpp", "tests") + os.sep in path:
            continue  # facade test programs are tests
        if pat.search(open(path, errors="ignore").read()):
            offenders.append(os.path.relpath(path, ROOT))
    assert not offenders, offenders


def test_package_imports_and_loads_its_library_with_the_oracle_unimportable():
    code = (
        "import sys\n"
        "sys.modules['oracle'] = None\n"          # any `import oracle` now raises ImportError
        "import bipedal_locomotion_framework_b200 as p\n"
        "from bipedal_locomotion_framework_b200 import _capi, contact_models, estimators, system, sharding, ini\n"
        "L = _capi.lib()\n"
        "assert b'sm_100a' in L.blf_ccm_version()\n"
        "print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_no_cpu_evaluation_path_in_the_python_harness():
    """Without a device the Python entry points raise / return False; they never compute."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    code = (
        "from bipedal_locomotion_framework_b200.contact_models import ContinuousContactModelBatch\n"
        "try:\n"
        "    ContinuousContactModelBatch(0)\n"
        "except Exception as e:\n"
        "    print('raised', type(e).__name__, str(e)[:200])\n"
        "else:\n"
        "    print('NO ERROR')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.stdout.startswith("raised"), r.stdout + r.stderr[-1000:]
    assert "no CPU path" in r.stdout or "CUDA" in r.stdout

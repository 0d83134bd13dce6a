"""The C++17 host layer (cpp/): same-named classes as the reference, computing through the C ABI.
Host-only parts run everywhere; the reference's three Catch2 sections need the GPU."""
import os
import subprocess

import pytest

from bipedal_locomotion_framework_b200 import build

LIB = build.LIB_DIR


@pytest.fixture(scope="module")
def cpp():
    build.build_cpp()
    return LIB


def _run(exe, *args):
    return subprocess.run([os.path.join(LIB, exe), *args], stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600)


def test_parameters_handler_cases(cpp):
    """src/ParametersHandler/tests/ParametersHandlerTest.cpp:25-117 against our StdImplementation."""
    r = _run("ParametersHandlerUnitTests")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failure(s)" in r.stdout


def test_host_expansion_helper_and_worker_pool(cpp):
    """csrc/host_expand.{h,cpp} on their own (no CUDA): every ISA form of the compact -> dense
    control-matrix expansion against a plain loop, bit for bit, for every destination alignment;
    the worker pool never re-runs the job of an earlier call after a resize."""
    r = _run("HostExpandUnitTests")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "All tests passed" in r.stdout


def test_facade_fails_loudly_without_a_gpu(cpp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run("ContinuousContactModelUnitTests", "Initialization")
    assert r.returncode != 0
    assert "there is no CPU evaluation path" in r.stderr
    # the handler-driven failures (missing key, wrong type, expired handler) are host logic and pass
    assert "Unable to get the variable named length" in r.stderr
    assert "The parameter handler is corrupted" in r.stderr


@pytest.mark.gpu
def test_reference_catch2_sections_on_the_gpu_facade(cpp):
    """ContinousContactModelTest.cpp:32-214 restated in cpp/tests/ContinuousContactModelTest.cpp,
    plus lazy-cache, batch-vs-instances and DeviceSoA / rollout arg-min cases."""
    r = _run("ContinuousContactModelUnitTests")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "4 test case(s)" in r.stdout and "0 failure(s)" in r.stdout
    print([ln for ln in r.stdout.splitlines() if "per-instance facade" in ln])


@pytest.mark.gpu
def test_reference_rls_test_on_the_gpu_estimator(cpp):
    """src/Estimators/tests/RecursiveLeastSquareTest.cpp:91-142 restated in
    cpp/tests/RecursiveLeastSquareTest.cpp against the GPU-backed estimator."""
    r = _run("RecursiveLeastSquareUnitTests")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "2 test case(s)" in r.stdout and "0 failure(s)" in r.stdout


def test_public_signatures_equal_the_reference(cpp):
    """cpp/tests/ApiConformanceTest.cpp: static_asserts spelling the reference's declarations
    (ContactModel, ContinuousContactModel, IParametersHandler, RecursiveLeastSquare, the System
    templates, ContactWrench) against the facade -- building it IS the test."""
    r = _run("ApiConformanceUnitTests")
    assert r.returncode == 0, r.stdout + r.stderr


def test_contact_wrench_holder(cpp):
    """System::ContactWrench (src/System/src/ContactWrench.cpp:13-35): frame index + shared model."""
    r = _run("ContactWrenchUnitTests")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failure(s)" in r.stdout


def test_generic_integrator_templates_host_section(cpp):
    """src/System/tests/IntegratorTest.cpp:27-78 (linear system through the generic ForwardEuler /
    FixedStepIntegrator templates; host logic only)."""
    r = _run("IntegratorUnitTests", "Linear")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "1 test case(s)" in r.stdout and "0 failure(s)" in r.stdout


@pytest.mark.gpu
def test_reference_integrator_test_and_batched_system_steps(cpp):
    """IntegratorTest.cpp:80-126 on the GPU-backed FloatingBaseSystemKinematics + ForwardEuler, and
    the batched Euler step / fused rollout / J^T wrench against loops over per-instance objects."""
    r = _run("IntegratorUnitTests")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "3 test case(s)" in r.stdout and "0 failure(s)" in r.stdout


@pytest.mark.gpu
def test_floating_base_dynamical_system_facade(cpp):
    """System::FloatingBaseDynamicalSystem (src/System/src/FloatingBaseSystemDynamics.cpp:17-251) over an
    injected KinDynComputations: every failure return of the reference, M acc = -h + sum J^T wrench + tau
    against the per-instance contact models, regularisation, no contacts / no joints, one ForwardEuler step."""
    r = _run("FloatingBaseSystemDynamicsUnitTests")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "2 test case(s)" in r.stdout and "0 failure(s)" in r.stdout

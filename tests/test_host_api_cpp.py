"""The C++ host layer (miden-vm_b200/host/miden_prover.hpp) mirrors the reference's Rust interface above the
C ABI: StarkConfig / Statement / ProverStatement / Preprocessed / ProverInstance::prove / ProverError.  Its parity
test is a C++ program (tests/cpp/test_host_api.cpp: prove -> verify -> tamper like the reference's own tests);
here it is built and run."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")


def _build():
    import __graft_entry__ as g
    g.build()
    subprocess.check_call(["make", "-s", "-C", CPP])
    return os.path.join(CPP, "test_host_api")


def test_cpp_host_api_builds_and_fails_loudly_without_a_device():
    import torch
    exe = _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "NO_DEVICE" in r.stdout and "no CPU fallback" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_host_api_prove_verify_tamper():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "HOST_API_OK" in r.stdout, r.stdout + r.stderr

"""The C ABI used from plain C (no Python, no torch in the loop): tests/c_abi/abi_smoke.c is compiled with gcc
against include/crb200.h + libcrb200.so and run on the GPU; it checks log|J|, x^T J^{-1} x, J^{-1} x and
diag(J^{-1}) against a dense Cholesky computed in the C program itself."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_c_client_of_the_abi(tmp_path):
    gcc = shutil.which("gcc")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if gcc is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("needs gcc and the CUDA runtime headers")
    libdir = os.path.join(ROOT, "cyclic-gps_b200")
    assert os.path.exists(os.path.join(libdir, "libcrb200.so")), "build libcrb200.so first (python cyclic-gps_b200/build.py)"
    exe = str(tmp_path / "abi_smoke")
    cmd = [gcc, "-std=c99", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c"), "-o", exe, "-L", libdir, "-lcrb200",
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_smoke ok" in r.stdout

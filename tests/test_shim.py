"""The C++ shim (include/pinocchio_b200_shim.hpp) compiles against the C ABI with plain g++ and behaves like the
reference's batched API: on a GPU box tests/cpp/shim_example.cpp checks aba(rnea(a)) == a and the
std::invalid_argument on a wrong-size argument; without a GPU it must fail loudly (no CPU fallback)."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


def _build(tmp, src="shim_example.cpp", extra=()):
    from pinocchio_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    exe = os.path.join(tmp, src[:-4])
    libdir = os.path.dirname(_capi.LIB_PATH)
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), *extra,
                           os.path.join(ROOT, "tests", "cpp", src), "-o", exe, "-L" + libdir,
                           "-lpinocchio_b200", "-Wl,-rpath," + libdir])
    return exe


EIGEN_STUB = ("-I" + os.path.join(ROOT, "tests", "cpp", "eigen_stub"),)


def test_eigen_overloads_compile_without_gpu(tmp_path):
    """Every Eigen-typed overload of the shim (the reference's parameter lists) compiles — against tests/cpp/eigen_stub, a
    stand-in for the members of Eigen::MatrixBase the shim uses, because Eigen is not in this image — and the program fails
    loudly without a GPU."""
    from pinocchio_b200 import _capi
    exe = _build(str(tmp_path), "shim_eigen_example.cpp", EIGEN_STUB)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    if _capi.lib().brbd_device_count() == 0:
        assert out.stdout.startswith("NO_GPU")


@pytest.mark.gpu
def test_eigen_overloads_run_on_gpu(tmp_path):
    exe = _build(str(tmp_path), "shim_eigen_example.cpp", EIGEN_STUB)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr


def test_shim_compiles_and_fails_loudly_without_gpu(tmp_path):
    from pinocchio_b200 import _capi
    exe = _build(str(tmp_path))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    if _capi.lib().brbd_device_count() == 0:
        assert out.stdout.startswith("NO_GPU") and "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_shim_runs_on_gpu(tmp_path):
    exe = _build(str(tmp_path))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr

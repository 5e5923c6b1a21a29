"""The header-only C++ adapter (PoseLib signatures over the C-ABI) compiles and links against the library with
stand-ins for the Eigen/PoseLib types; with a GPU it also runs every entry point once."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_adapter_test")


def _build():
    import __graft_entry__ as ge
    ge.build()
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "adapter_compile_test.cc"),
                           "-o", EXE, "-L" + os.path.join(ROOT, "poselib_b200"), "-lposelib_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "poselib_b200")])


def test_adapter_compiles_and_links():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "adapter link ok" in out.stdout, out.stderr


@pytest.mark.gpu
def test_adapter_runs_every_entry_point():
    _build()
    out = subprocess.run([EXE, "run"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "adapter run ok" in out.stdout, out.stdout + out.stderr

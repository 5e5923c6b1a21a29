"""The drop-in translation unit (poselib_b200/adapter/poselib_dropin.cc) is compiled against the REFERENCE'S OWN headers
(PoseLib/robust.h, robust/ransac.h, robust/bundle.h, solvers/*.h) and linked, together with libposelib_b200.so, into a
client that includes PoseLib's headers only.  Its 18 definitions must match PoseLib's declarations exactly, otherwise the
client does not link.  Eigen3 is not installed here; the Eigen stand-in of the oracle tree (oracle/ref/mini) is used at
compile time only.  Needs /root/reference (CPU container); the compiled client also has a `run` mode for a GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
EXE = os.path.join(ROOT, "tests", "_dropin_client")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "PoseLib")), reason="needs the reference's headers")


def test_dropin_defines_poselibs_own_declarations():
    import __graft_entry__ as ge
    ge.build()
    inc = ["-I" + os.path.join(ROOT, "oracle", "ref", "mini"), "-I" + os.path.join(ROOT, "oracle"), "-I" + REF,
           "-I" + os.path.join(ROOT, "poselib_b200", "adapter")]
    subprocess.check_call(
        ["g++", "-std=c++17", "-O1", "-w", *inc,
         os.path.join(ROOT, "tests", "dropin_client_test.cc"),
         os.path.join(ROOT, "poselib_b200", "adapter", "poselib_dropin.cc"),
         os.path.join(REF, "PoseLib", "misc", "camera_models.cc"),  # Camera's constructors: stays PoseLib's own code
         "-o", EXE, "-L" + os.path.join(ROOT, "poselib_b200"), "-lposelib_b200",
         "-Wl,-rpath," + os.path.join(ROOT, "poselib_b200")])
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "dropin link ok: 18 PoseLib entry points resolved" in out.stdout, out.stdout + out.stderr

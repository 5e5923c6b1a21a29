import os
import sys


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    # Several test modules import the C-ABI binding at collection time; it refuses to load without the in-tree library.
    # A fresh checkout that runs pytest before __graft_entry__.build() would otherwise fail to collect.
    if not (os.path.exists(os.path.join(ROOT, "poselib_b200", "libposelib_b200.so"))
            and os.path.exists(os.path.join(ROOT, "oracle", "_build", "libplo.so"))):
        import __graft_entry__ as ge
        ge.build()

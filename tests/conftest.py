import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _cap_cpu_threads():
    # ATen's small convolutions get slower, not faster, with more threads than the container can schedule
    try:
        import torch
        torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    except Exception:
        pass


_cap_cpu_threads()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))

    return load


def _ensure_native_built():
    """The .so is git-ignored; build it in-tree if this checkout does not have it yet (nvcc
    cross-compiles without a GPU).  On the GPU box the prebuilt .so travels with the snapshot."""
    so = os.path.join(ROOT, "asv_subtools_b200", "libxvb200.so")
    exe = os.path.join(ROOT, "asv_subtools_b200", "bin", "xvb-extract")
    if not (os.path.exists(so) and os.path.exists(exe)):
        import shutil
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            import __graft_entry__
            __graft_entry__.build()


_ensure_native_built()

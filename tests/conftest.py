import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def hostmath():
    """ctypes handle of the TEST-ONLY host build of csrc/fs_math.cuh (tests/hostmath/hostmath.cpp)."""
    import ctypes

    src = ROOT / "tests" / "hostmath" / "hostmath.cpp"
    hdr = ROOT / "fusionsense_b200" / "csrc" / "fs_math.cuh"
    so = ROOT / "tests" / "_hostmath.so"
    if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", str(so), str(src)])
    return ctypes.CDLL(str(so))

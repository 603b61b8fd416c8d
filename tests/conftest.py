import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def hostemu():
    """tests/hostemu/libhostemu.so: the product's per-surfel __host__ __device__ math compiled for the CPU."""
    import ctypes
    d = os.path.join(ROOT, "tests", "hostemu")
    so = os.path.join(d, "libhostemu.so")
    src = os.path.join(d, "hostemu.cpp")
    deps = [src, os.path.join(ROOT, "eggfusion_b200", "csrc", "egs_surfel_math.cuh"),
            os.path.join(ROOT, "eggfusion_b200", "csrc", "egs_common.cuh"),
            os.path.join(ROOT, "eggfusion_b200", "csrc", "egm_math.cuh"),
            os.path.join(ROOT, "eggfusion_b200", "csrc", "egt_gn_math.cuh"),
            os.path.join(ROOT, "eggfusion_b200", "csrc", "egt_qr.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(p) for p in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-I/usr/local/cuda/include", "-shared",
                        "-o", so, src], check=True)
    return ctypes.CDLL(so)

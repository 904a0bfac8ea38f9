import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False
    n = ctypes.c_int(0)
    return cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a GPU should say so instead of failing inside CUDA
    if any("gpu" in item.keywords for item in items) and not _have_gpu():
        skip = pytest.mark.skip(reason="no CUDA device")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


@pytest.fixture(scope="session")
def hostemul():
    """tests/hostemul: the product's per-pixel device routines compiled for the CPU (stride 1)."""
    d = os.path.join(ROOT, "tests", "hostemul")
    so = os.path.join(d, "libnl_hostemul.so")
    srcs = [os.path.join(d, "host_emul.cpp"), os.path.join(ROOT, "nightlight_b200", "csrc", "nl_column.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-std=c++17",
                               "-o", so, srcs[0]])
    import ctypes as C
    L = C.CDLL(so)
    fp = C.POINTER(C.c_float)
    L.emul_stack.restype = C.c_int
    L.emul_stack.argtypes = [C.c_int, C.POINTER(fp), C.c_int, C.c_size_t, fp, C.c_float, C.c_float, C.c_float, fp,
                             C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.emul_qselect_median.restype = C.c_float
    L.emul_qselect_median.argtypes = [fp, C.c_int]
    L.emul_sort.argtypes = [fp, C.c_int, C.c_int]
    return L


@pytest.fixture(scope="session")
def ctx():
    import nightlight_b200 as nl
    c = nl.Context(0)
    yield c
    c.close()


@pytest.fixture
def tuning(ctx):
    """nl_ctx_set_tuning for one test; the session context goes back to the built-in settings afterwards"""
    used = []

    def set_(key, value):
        used.append(key)
        ctx.set_tuning(key, value)

    yield set_
    for key in used:
        ctx.set_tuning(key, {"defer_passes": "", "linfit_stream": "1"}.get(key, "0"))

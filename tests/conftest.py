import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with `-m gpu`; without a device they are skipped, not failed.
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


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "re2-modification_b200")


def hostsim_lib_path():
    return os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")


@pytest.fixture(scope="session", autouse=True)
def built_checkers(request):
    """CPU tier: the C restatement (oracle/) and tests/hostsim/libhostsim.so -- the kernels' cores and the
    kernel SOURCES compiled for the host under the SIMT emulator -- rebuilt when a source is newer."""
    import subprocess
    if (request.config.getoption("-m") or "").strip() == "gpu":
        return  # the GPU tier does not use the emulator
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, stdout=subprocess.DEVNULL)
    so = hostsim_lib_path()
    src = [os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp"),
           os.path.join(ROOT, "tests", "hostsim", "kernels_simt.cpp"),
           os.path.join(PKG, "csrc", "rxm_plan.cpp")]
    deps = src + [os.path.join(PKG, "csrc", f) for f in os.listdir(os.path.join(PKG, "csrc"))
                  if f.endswith((".cu", ".cuh", ".hpp"))] + [os.path.join(ROOT, "tests", "hostsim", "simt_shim.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-x", "c++",
                        "-o", so, *src], check=True)

import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


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
def orc():
    from oracle import slfp_oracle
    slfp_oracle.lib()
    return slfp_oracle


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def g_quant():
    return _load("quant_samples.npz")


@pytest.fixture(scope="session")
def g_tables():
    return _load("quant_tables.npz")


@pytest.fixture(scope="session")
def g_conv():
    return _load("conv_cases.npz")


@pytest.fixture(scope="session")
def g_act():
    return _load("act_cases.npz")


@pytest.fixture(scope="session")
def g_sgd():
    return _load("sgd_cases.npz")


@pytest.fixture(scope="session")
def hostcheck():
    """The product's encode/decode device functions compiled for the host (tests/host_check.cu)."""
    so = os.path.join(ROOT, "tests", "_build", "libhostcheck.so")
    src = os.path.join(ROOT, "tests", "host_check.cu")
    dep = os.path.join(ROOT, "cnns_slfp_quantization_b200", "csrc", "slfp_common.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                               "-Wno-deprecated-gpu-targets", src, "-o", so])
    return ctypes.CDLL(so)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_bits(ref, got):
    ref, got = np.asarray(ref, np.float32), np.asarray(got, np.float32)
    eq = bits(ref) == bits(got)
    return eq | (np.isnan(ref) & np.isnan(got))

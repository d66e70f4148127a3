import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def unpack(npz, prefix=""):
    d = {}
    for split in ("train", "val"):
        n = int(npz[f"{prefix}{split}_n"])
        d[split] = [{k: npz[f"{prefix}{split}{i}_{k}"] for k in ("t", "y", "u")} for i in range(n)]
    return d


@pytest.fixture(scope="session")
def arm_data():
    return unpack(np.load(os.path.join(GOLDEN, "arm_data.npz")))


@pytest.fixture(scope="session")
def snake_data():
    return unpack(np.load(os.path.join(GOLDEN, "snake_data.npz")))


@pytest.fixture(scope="session")
def rsys_data():
    z = np.load(os.path.join(GOLDEN, "rsys_subset.npz"))
    return [unpack(z, prefix=f"s{i}_") for i in range(int(z["nsys"]))]


@pytest.fixture(scope="session")
def golden_Z():
    return np.load(os.path.join(GOLDEN, "arm_blockM_Z.npz"))


@pytest.fixture(scope="session")
def hostlift():
    """TEST-ONLY host evaluator of the dictionary (g++ build of program.cpp + lift_eval.h)."""
    import ctypes as C
    so = os.path.join(ROOT, "tests", "_build", "hostlift.so")
    src = [os.path.join(ROOT, "tests", "hostlift.cpp"), os.path.join(ROOT, "koopman-realizations_b200", "csrc", "program.cpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", so] + src)
    return C.CDLL(so)


@pytest.fixture(scope="session")
def fitter():
    import koopfit
    f = koopfit.Fitter(device=0)
    yield f
    f.close()

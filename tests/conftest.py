import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

PKG = "engineering-degree-in-plasma-simulations_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs oracle/_ref (the compiled reference); skipped when absent")


@pytest.fixture(scope="session")
def picgpu():
    """The product binding (ctypes over libpicgpu.so), initialised on cuda:0."""
    mod = importlib.import_module(PKG + ".picgpu")
    mod.init(0)
    return mod


@pytest.fixture(scope="session")
def orc():
    from oracle import pic_oracle
    pic_oracle.lib()
    return pic_oracle


@pytest.fixture(scope="session")
def ref():
    from oracle import ref_v3
    if not ref_v3.available():
        pytest.skip("oracle/_ref/libref_v3.so not built (reference tree absent)")
    ref_v3.lib()
    ref_v3.config(subcycling=False, multithreading=False, merging=False, sputtering=False)
    return ref_v3


@pytest.fixture(scope="session")
def ref2():
    """The compiled ch4/v2 reference (BASELINE config 3: fixed-weight MC_MEX_Ionization)."""
    from oracle import ref_ch4v2
    if not ref_ch4v2.available():
        pytest.skip("oracle/_ref/libref_ch4v2.so not built (reference tree absent)")
    ref_ch4v2.lib()
    return ref_ch4v2

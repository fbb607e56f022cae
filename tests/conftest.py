import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def s3d():
    """The product package (its directory name starts with a digit)."""
    mod = importlib.import_module("3dsift_b200")
    mod.lib()
    return mod


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("3dsift_b200.synth")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import ref as O
    return O


@pytest.fixture(scope="session")
def port(oracle_mod):
    """The C restatement (always buildable: gcc only)."""
    return oracle_mod.Port()


@pytest.fixture(scope="session")
def refimpl(oracle_mod):
    """The compiled reference (oracle/_ref) — prebuilt where /root/reference exists; travels to
    the GPU box with the snapshot."""
    if not oracle_mod.have_ref():
        if os.path.isdir("/root/reference/3DSIFT"):
            oracle_mod.build("ref")
        else:
            pytest.skip("oracle/_ref/libsift3d_ref.so not present and /root/reference absent")
    return oracle_mod.Ref()


@pytest.fixture(scope="session")
def checker(oracle_mod):
    """Strongest checker available (compiled reference, else the port)."""
    if not oracle_mod.have_ref() and os.path.isdir("/root/reference/3DSIFT"):
        oracle_mod.build("ref")
    return oracle_mod.best()


def shell_mask(shape):
    """True on the outer 1-voxel shell of a [z,y,x] volume."""
    m = np.ones(shape, bool)
    m[1:-1, 1:-1, 1:-1] = False
    return m

import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pkg(name=""):
    """import 3dfeatnet_b200[.name] (a package name starting with a digit needs importlib)."""
    return importlib.import_module("3dfeatnet_b200" + ("." + name if name else ""))


@pytest.fixture(scope="session")
def oracle_ops():
    from oracle import ops

    ops.lib()
    return ops


@pytest.fixture(scope="session")
def f3d_lib():
    """The C-ABI library; built in-tree if nvcc is present and the .so is missing."""
    lib_mod = pkg("_lib")
    if not os.path.exists(lib_mod.LIB_PATH):
        pkg("build").build()
    return lib_mod.lib()


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    pkg("_lib").lib()
    return torch.device("cuda:0")

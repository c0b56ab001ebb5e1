"""CPU test of the drop-in boundary: the shared library loads without a GPU and
exports every symbol include/bayeslogit_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bayeslogit_b200.h")

# the reference's exported C symbols, Code/C/LogitWrapper.h:23-64
REFERENCE_SYMBOLS = ["rpg_gamma", "rpg_devroye", "rpg_alt", "rpg_sp", "rpg_hybrid",
                     "gibbs", "EM", "combine", "mult_gibbs", "mult_combine"]


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(\w+)\s*\([^;{]*\)\s*;", text, flags=re.M)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
    from bayeslogit_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build_native()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_declares_reference_symbols():
    names = declared_symbols()
    for s in REFERENCE_SYMBOLS:
        assert s in names
    assert len(names) > 30


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_version_and_loud_failure_without_gpu(lib):
    import torch
    lib.bl_version.restype = ctypes.c_int
    assert lib.bl_version() >= 100
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is not reachable")
    # no CPU fallback: a compute call without a device must report an error
    import numpy as np
    from bayeslogit_b200 import _lib, api
    with pytest.raises(_lib.EngineError):
        api.rpg_devroye(4, 1, np.zeros(4))

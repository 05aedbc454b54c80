"""The C-ABI library loads without a GPU and exports every symbol include/vgm_b200.h declares;
the product fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vgm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vgm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from variational_gradient_matching_for_dynamical_systems_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH), "build with __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    # and the ctypes binding covers all of them with explicit signatures
    assert set(names) <= set(_lib.SIGNATURES), set(names) - set(_lib.SIGNATURES)
    assert _lib.load().vgm_version() >= 100


def test_struct_sizes_match_c_layout():
    """Field offsets of the ctypes mirrors follow the C structs (natural alignment, x86-64)."""
    from variational_gradient_matching_for_dynamical_systems_b200 import _lib
    assert ctypes.sizeof(_lib.GemmArgs) == 3 * 8 + 6 * 4 + 3 * 8 + 2 * 8 + 8 + 3 * 8 + 3 * 8 + 2 * 8 + 8
    assert ctypes.sizeof(_lib.SolverOpts) == 16
    assert ctypes.sizeof(_lib.SweepInfo) == 16
    assert ctypes.sizeof(_lib.Operators) == 40


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from variational_gradient_matching_for_dynamical_systems_b200 import OdeModel
    from variational_gradient_matching_for_dynamical_systems_b200.engine import VGMEngine
    import numpy as np
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VGMEngine(OdeModel.lorenz96(10), np.eye(8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "variational_gradient_matching_for_dynamical_systems_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f

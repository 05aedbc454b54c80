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


def test_struct_sizes_match_c_layout(tmp_path):
    """The ctypes mirrors have the layout gcc gives the C structs of include/vgm_b200.h."""
    import subprocess
    from variational_gradient_matching_for_dynamical_systems_b200 import _lib
    pairs = {"vgm_gemm_args": _lib.GemmArgs, "vgm_solver_opts": _lib.SolverOpts, "vgm_sweep_info": _lib.SweepInfo,
             "vgm_operators": _lib.Operators, "vgm_sweep_plan": _lib.SweepPlan, "vgm_model_tables": _lib.ModelTables}
    lines = []
    for cname, ct in pairs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vgm_b200.h"\nint main(void){%s return 0;}\n'
                   % "".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, ct in pairs.items():
        assert int(out[cname]) == ctypes.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out["%s.%s" % (cname, fname)]) == getattr(ct, fname).offset, (cname, fname)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from variational_gradient_matching_for_dynamical_systems_b200 import OdeModel
    from variational_gradient_matching_for_dynamical_systems_b200.engine import VGMEngine
    import numpy as np
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VGMEngine(OdeModel.lorenz96(10), np.eye(8))


def test_comm_wrappers_resolve_nccl_at_run_time():
    """vgm_comm_*: NCCL is dlopen'ed on first use (the copy PyTorch already loaded), the id needs no GPU, and a
    communicator cannot be created without a device (no CPU fallback)."""
    import torch
    from variational_gradient_matching_for_dynamical_systems_b200 import _lib
    lib = _lib.load()
    ver = ctypes.c_int(0)
    assert lib.vgm_comm_nccl_version(ctypes.byref(ver)) == _lib.VGM_OK, lib.vgm_last_error()
    major, minor, patch = torch.cuda.nccl.version()[:3]
    assert ver.value == major * 10000 + minor * 100 + patch
    assert lib.vgm_comm_halo_ws_bytes(8, 6, 1024) == 9 * 6 * 1024 * 8
    if not torch.cuda.is_available():
        ident = ctypes.create_string_buffer(_lib.COMM_ID_BYTES)
        comm = ctypes.c_void_p()
        assert lib.vgm_comm_init(ident.raw, 0, 1, ctypes.byref(comm)) in (_lib.VGM_ECUDA, _lib.VGM_ENCCL)
        assert not comm.value


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "variational_gradient_matching_for_dynamical_systems_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f

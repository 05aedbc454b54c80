"""Stand-alone timing of the banded DMMA GEMM at the shape of the C5 residual (N = 50 000 rows, M = K = 1000, ld 1024,
CUDA events; A and the output are 410 MB each, larger than L2).  TFLOP/s are quoted for the DENSE-EQUIVALENT band
work 2 N M (2 band + 1) so that tile variants and the ragged band compare by time."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from variational_gradient_matching_for_dynamical_systems_b200 import _lib  # noqa: E402

lib = _lib.load()
N, T, ld = 50000, 1000, 1024
X = torch.randn(N, ld, device="cuda", dtype=torch.float64)
B = torch.randn(T, ld, device="cuda", dtype=torch.float64)
out = torch.empty(N, ld, device="cuda", dtype=torch.float64)
variants = [int(v) for v in os.environ.get("VARIANTS", "12,2,0").split(",")]
for variant in variants:
    lib.vgm_dgemm_set_variant(variant)
    for band in (126, 119, 0):
        g = _lib.GemmArgs()
        g.A, g.B, g.C = X.data_ptr(), B.data_ptr(), out.data_ptr()
        g.lda = g.ldb = g.ldc = ld
        g.N, g.M, g.K = N, T, T
        g.alpha = 1.0
        g.band = band
        for _ in range(2):
            _lib.check(lib.vgm_dgemm_nt(C.byref(g), _lib.stream_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            lib.vgm_dgemm_nt(C.byref(g), _lib.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 200
        flops = 2.0 * N * T * ((2 * band + 1) if band else T)
        print("variant %2d band %3d: %8.1f us   %.1f TFLOP/s (band-only work)" % (variant, band, us, flops / us / 1e6))

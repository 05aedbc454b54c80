"""Stand-alone timing of the tcgen05 GEMM at the shapes of the C5 correction solve (CUDA events, L2 flushed by
the operand sizes: 100-200 MB of A and 200 MB of output per launch)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from variational_gradient_matching_for_dynamical_systems_b200 import _lib  # noqa: E402

lib = _lib.load()
T = 1000
ld = 1000
M = torch.randn(2, T, ld, device="cuda").to(torch.bfloat16)
for n in (50000, 25000, 6250):
    P = torch.randn(2, n, ld, device="cuda").to(torch.bfloat16)
    out = torch.empty(n, ld, device="cuda", dtype=torch.float32)
    for split, band in ((True, 146), (True, 47), (False, 133), (False, 36), (True, 0), (False, 0)):
        args = (P[0].data_ptr(), P[1].data_ptr() if split else None, n, ld, M[0].data_ptr(),
                M[1].data_ptr() if split else None, T, T, ld, band, out.data_ptr(), ld, None)
        for _ in range(3):
            _lib.check(lib.vgm_tc_gemm_bf16(*args))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.vgm_tc_gemm_bf16(*args)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        kb = sum(-(-min(T, m0 + 128 + band) // 64) - max(0, m0 - band) // 64 if band else 16 for m0 in range(0, T, 128))
        tiles = -(-n // 128)
        l2 = tiles * kb * (65536 if split else 32768)
        hbm = n * T * (4 if split else 2) + n * T * 4
        print("n=%6d split=%d band=%3d: %7.1f us  L2->SMEM %.2f TB/s  HBM %.2f TB/s  MMA %.0f TFLOP/s" %
              (n, split, band, us, l2 / us / 1e6, hbm / us / 1e6, 2.0 * tiles * 128 * 128 * kb * 64 * (3 if split else 1) / us / 1e6))

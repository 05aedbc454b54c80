"""Standalone check + timing of the tcgen05 BF16 GEMM (run under `timeout`)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from variational_gradient_matching_for_dynamical_systems_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = "cuda"
torch.manual_seed(0)


def run(N, M, K, n_act=None, gather=False, dots=True):
    ld = (K + 7) // 8 * 8
    ldz = (M + 7) // 8 * 8
    A = torch.zeros((N, ld), dtype=torch.bfloat16, device=dev)
    B = torch.zeros((M, ld), dtype=torch.bfloat16, device=dev)
    A[:, :K] = torch.randn((N, K), device=dev).to(torch.bfloat16)
    B[:, :K] = torch.randn((M, K), device=dev).to(torch.bfloat16)
    n_act = N if n_act is None else n_act
    rows = torch.randperm(N, device=dev)[:n_act].to(torch.int32).contiguous() if gather else None
    Z = torch.full((N, ldz), -7.0, dtype=torch.float64, device=dev)
    W = torch.randn((N, ldz), dtype=torch.float64, device=dev)
    nt = (M + 255) // 256
    dots_out = torch.zeros((N, nt), dtype=torch.float64, device=dev)
    nrd = torch.tensor([n_act], dtype=torch.int32, device=dev)
    args = (A.data_ptr(), N, ld, B.data_ptr(), M, K, ld, n_act, nrd.data_ptr(), None if rows is None else rows.data_ptr(),
            Z.data_ptr(), ldz, W.data_ptr() if dots else None, dots_out.data_ptr() if dots else None, nt,
            _lib.stream_ptr())
    _lib.check(lib.vgm_tc_gemm_bf16(*args))
    torch.cuda.synchronize()
    ref = (A[:n_act, :K].float() @ B[:, :K].float().t()).double()
    phys = rows.long() if rows is not None else torch.arange(n_act, device=dev)
    got = Z[phys, :M]
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    derr = 0.0
    if dots:
        dref = (W[phys, :M] * ref).sum(1)
        derr = (dots_out[phys].sum(1) - dref).abs().max().item() / dref.abs().max().item()
    untouched = torch.ones(N, dtype=torch.bool, device=dev)
    untouched[phys] = False
    clean = bool((Z[untouched] == -7.0).all().item())
    # timing
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        lib.vgm_tc_gemm_bf16(*args)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        lib.vgm_tc_gemm_bf16(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("N=%d M=%d K=%d n_act=%d gather=%s: relerr %.2e  dot relerr %.2e  untouched rows clean=%s  %.3f ms  %.1f TFLOP/s"
          % (N, M, K, n_act, gather, err, derr, clean, ms, 2.0 * n_act * M * K / ms / 1e9), flush=True)
    assert err < 2e-3 and derr < 1e-5 and clean


run(128, 256, 64)
run(128, 256, 256)
run(300, 1000, 1000)
run(1000, 99, 99)
run(700, 1000, 1000, n_act=333, gather=True)
run(50000, 1000, 1000)
run(50000, 1000, 1000, dots=False)
print("tc_gemm ok")

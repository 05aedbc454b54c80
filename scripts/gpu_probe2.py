"""GEMM tile-variant sweep + BF16-vs-FP64 preconditioner comparison.  Writes gpurun_out/probe2.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from variational_gradient_matching_for_dynamical_systems_b200 import OdeModel, _lib  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.engine import Operators, VGMEngine  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.GP_regression import kernel_function_device  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.synthetic import lorenz96_problem  # noqa: E402

out = {}
dev = "cuda"
lib = _lib.load()


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


b = torch.randn((1000, 1000), dtype=torch.float64, device=dev)
for N in (50000, 6250):
    A = torch.randn((N, 1000), dtype=torch.float64, device=dev)
    Cm = torch.zeros((N, 1000), dtype=torch.float64, device=dev)
    ref = A @ b.t()
    for v in range(5):
        lib.vgm_dgemm_set_variant(v)
        g = _lib.GemmArgs()
        g.A, g.B, g.C = A.data_ptr(), b.data_ptr(), Cm.data_ptr()
        g.lda = g.ldb = g.ldc = 1000
        g.N, g.M, g.K = N, 1000, 1000
        g.alpha = 1.0
        s = _lib.stream_ptr()
        try:
            Cm.zero_()
            best = timeit(lambda: _lib.check(lib.vgm_dgemm_nt(C.byref(g), s)))
            err = (Cm - ref).abs().max().item()
            out["dgemm_N%d_v%d" % (N, v)] = {"ms": best, "tflops": 2 * N * 1e6 / best / 1e9, "maxerr": err}
        except Exception as e:  # noqa: BLE001
            out["dgemm_N%d_v%d" % (N, v)] = repr(e)
        print(N, v, out["dgemm_N%d_v%d" % (N, v)], flush=True)
    del A, Cm, ref
lib.vgm_dgemm_set_variant(0)

T = 1000
tg = np.arange(T) * 0.05
kf = kernel_function_device(tg, tg, "rbf", [0.0625, 1.0])
ops = Operators(kf["D"][:, :T], None)
for K in (1000, 100000):
    _, X0, hidden = lorenz96_problem(K, T, device=dev)
    model = OdeModel.lorenz96(K)
    for dt in ("bf16", "f64"):
        eng = VGMEngine(model, ops, None, hidden=hidden, precond_dtype=dt)
        times = []
        for it in range(3):
            eng.set_states_state_major(X0)
            torch.cuda.synchronize()
            t0 = time.time()
            info = eng.iteration(allow_unconverged=True)
            torch.cuda.synchronize()
            times.append(time.time() - t0)
        Xres = eng.X[:, :T].clone()
        out["K%d_%s" % (K, dt)] = {"iter_s": times, "info": info, "updates_per_s": len(hidden) * T / min(times)}
        if dt == "bf16":
            Xb = Xres
        else:
            out["K%d_bf16_vs_f64_state_relerr" % K] = ((Xb - Xres).abs().max() / Xres.abs().max()).item()
        print(K, dt, out["K%d_%s" % (K, dt)], flush=True)
        del eng
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe2.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

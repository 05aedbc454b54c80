"""One-off GPU probe: FP64 GEMM peaks (cuBLAS via torch vs ours) and first end-to-end timings.
Writes gpurun_out/probe.json."""
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
from variational_gradient_matching_for_dynamical_systems_b200.engine import VGMEngine  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.synthetic import lorenz96_problem  # noqa: E402
from oracle import vgm_oracle as vo  # noqa: E402

out = {}
dev = "cuda"


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


# ---- cuBLAS FP64 peak -----------------------------------------------------------------
for n in (4096, 8192):
    a = torch.randn((n, n), dtype=torch.float64, device=dev)
    b = torch.randn((n, n), dtype=torch.float64, device=dev)
    best, med = timeit(lambda: torch.matmul(a, b), n=5)
    out["cublas_dgemm_%d_tflops" % n] = 2 * n ** 3 / best / 1e9
    del a, b
a = torch.randn((50000, 1000), dtype=torch.float64, device=dev)
b = torch.randn((1000, 1000), dtype=torch.float64, device=dev)
best, med = timeit(lambda: torch.matmul(a, b.t()), n=5)
out["cublas_dgemm_50000x1000x1000_tflops"] = 2 * 50000 * 1e6 / best / 1e9
print(out, flush=True)

# ---- our DMMA GEMM -------------------------------------------------------------------------
lib = _lib.load()
for N in (50000, 6250, 500, 16, 1):
    A = torch.randn((N, 1000), dtype=torch.float64, device=dev)
    Cm = torch.zeros((N, 1000), dtype=torch.float64, device=dev)
    g = _lib.GemmArgs()
    g.A, g.B, g.C = A.data_ptr(), b.data_ptr(), Cm.data_ptr()
    g.lda = g.ldb = g.ldc = 1000
    g.N, g.M, g.K = N, 1000, 1000
    g.alpha = 1.0
    s = _lib.stream_ptr()
    best, med = timeit(lambda: _lib.check(lib.vgm_dgemm_nt(C.byref(g), s)), n=5)
    out["vgm_dgemm_%dx1000x1000_tflops" % N] = 2 * N * 1e6 / best / 1e9
    out["vgm_dgemm_%dx1000x1000_ms" % N] = best
    err = (Cm - A @ b.t()).abs().max().item()
    out["vgm_dgemm_%d_maxerr" % N] = err
    del A, Cm
print(out, flush=True)

# ---- first end-to-end timings ------------------------------------------------------------------
T = 1000
t = np.arange(T) * 0.05
D, A_, *_ = vo.kernel_function(t, t, "rbf", [0.0625, 1.0])
for K in (1000, 12500, 100000):
    try:
        t0 = time.time()
        _, X0, hidden = lorenz96_problem(K, T, device=dev)
        t_data = time.time() - t0
        t0 = time.time()
        model = OdeModel.lorenz96(K)
        eng = VGMEngine(model, D, None, hidden=hidden)
        eng.set_states_state_major(X0)
        torch.cuda.synchronize()
        t_setup = time.time() - t0
        res = {"t_data_s": t_data, "t_setup_s": t_setup}
        for precond in (True, False):
            eng.precondition = precond
            eng._cops.G = None
            times, infos = [], []
            for it in range(3):
                eng.set_states_state_major(X0)
                torch.cuda.synchronize()
                t0 = time.time()
                info = eng.iteration(allow_unconverged=True)
                torch.cuda.synchronize()
                times.append(time.time() - t0)
                infos.append(info)
            key = "precond" if precond else "plain"
            res[key + "_iter_s"] = times
            res[key + "_info"] = infos[-1]
            res[key + "_updates_per_s"] = len(hidden) * T / min(times)
        res["theta"] = eng.get_theta().tolist()
        res["shift"] = eng.ops.G_shift
        out["lorenz96_K%d" % K] = res
        print(K, res, flush=True)
        del eng
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["lorenz96_K%d_error" % K] = repr(e)
        print("ERROR", K, repr(e), flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))

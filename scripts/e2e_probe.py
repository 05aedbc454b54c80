"""e2e probe: vgm_iteration_host with / without the strided copy-out of the updated rows, with / without pipelining."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from variational_gradient_matching_for_dynamical_systems_b200 import OdeModel  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.engine import Operators, VGMEngine  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.GP_regression import kernel_function_device  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.synthetic import lorenz96_problem  # noqa: E402


def main():
    K, T = 100000, 1000
    tg = np.arange(T) * 0.05
    kf = kernel_function_device(tg, tg, "rbf", [0.0625, 1.0])
    ops = Operators(kf["D"][:, :T], None)
    _, X0, hidden = lorenz96_problem(K, T, 0.05, seed=0)
    eng = VGMEngine(OdeModel.lorenz96(K), ops, None, hidden=hidden)
    eng.set_states_state_major(X0)
    X0l = eng.X.clone()
    Xin = torch.empty((K, eng.ld), dtype=torch.float64, pin_memory=True)
    Xin.copy_(X0l)
    Xout = torch.empty((K, eng.ld), dtype=torch.float64, pin_memory=True)
    Xout.copy_(X0l)
    th = torch.zeros(1, dtype=torch.float64, pin_memory=True)
    out = {}
    for name, rows_only, chunks in (("rows_only_pipe4", True, (0.12, 0.5, 0.88, 1.0)), ("full_out_pipe4", False, (0.12, 0.5, 0.88, 1.0)),
                                    ("rows_only_nopipe", True, 0), ("full_out_nopipe", False, 0)):
        for _ in range(2):
            eng.iteration_host(Xin, th, Xout, updated_rows_only=rows_only, pipeline_chunks=chunks)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            eng.iteration_host(Xin, th, Xout, updated_rows_only=rows_only, pipeline_chunks=chunks)
        torch.cuda.synchronize()
        out[name] = round((time.perf_counter() - t0) / 4 * 1e3, 2)
    # raw copies
    dst = torch.empty_like(X0l)
    for name, fn in (("h2d_800MB", lambda: dst.copy_(Xin, non_blocking=True)), ("d2h_800MB", lambda: Xout.copy_(dst, non_blocking=True))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        out[name + "_ms"] = round((time.perf_counter() - t0) / 3 * 1e3, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Does a host<->device copy running beside the solver slow the solver down?  (context for the e2e figure:
vgm_iteration_host overlaps 800 MB of H2D and 400 MB of D2H with a step that is bound by HBM traffic)"""
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
    host = torch.empty((K, eng.ld), dtype=torch.float64, pin_memory=True)
    host.copy_(X0l)
    dst = torch.empty_like(X0l)
    back = torch.empty((K // 2, eng.ld), dtype=torch.float64, pin_memory=True)
    side_in, side_out = torch.cuda.Stream(), torch.cuda.Stream()

    def step_ms(n=4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            eng.X.copy_(X0l)
            eng.iteration()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for _ in range(2):
        step_ms(1)
    out = {"alone_ms": step_ms()}
    for name, h2d, d2h in (("with_h2d", True, False), ("with_d2h", False, True), ("with_both", True, True)):
        torch.cuda.synchronize()
        # enough copies queued to cover the timed steps (each H2D of 819 MB is ~15 ms)
        if h2d:
            with torch.cuda.stream(side_in):
                for _ in range(14):
                    dst.copy_(host, non_blocking=True)
        if d2h:
            with torch.cuda.stream(side_out):
                for _ in range(28):
                    back.copy_(dst[:K // 2], non_blocking=True)
        t0 = time.perf_counter()
        out[name + "_ms"] = step_ms()
        torch.cuda.synchronize()
        out[name + "_copies_s"] = time.perf_counter() - t0
    print(json.dumps(out))


if __name__ == "__main__":
    main()

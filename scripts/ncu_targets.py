"""Workload for the ncu captures (run under `ncu ... python scripts/ncu_targets.py [--K 100000]`): one
kernel_function set-up at T = 1000 (K1 kernel_build, K2 potrf/trsm + DGEMM) and two full coordinate-ascent iterations
of Lorenz96 (K4 param_stats, K5 state_assemble, the solver kernels) -- the second iteration is the steady state."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from variational_gradient_matching_for_dynamical_systems_b200 import OdeModel  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.engine import Operators, VGMEngine  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.GP_regression import kernel_function_device  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.synthetic import lorenz96_problem  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=100000)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    nvtx = torch.cuda.nvtx
    tg = np.arange(a.T) * 0.05
    nvtx.range_push("setup")             # ncu --nvtx --nvtx-include "setup/" : kernel_build, Cholesky, operators
    kf = kernel_function_device(tg, tg, "rbf", [0.0625, 1.0])
    ops = Operators(kf["D"][:, :a.T], None)
    torch.cuda.synchronize()
    nvtx.range_pop()
    _, X0, hidden = lorenz96_problem(a.K, a.T, 0.05, seed=0)
    eng = VGMEngine(OdeModel.lorenz96(a.K), ops, None, hidden=hidden)
    eng.set_states_state_major(X0)
    X0l = eng.X.clone()
    for it in range(a.iters):
        eng.X.copy_(X0l)
        torch.cuda.synchronize()
        nvtx.range_push("steady" if it == a.iters - 1 else "warm")     # --nvtx-include "steady/": the last iteration
        eng.iteration()
        torch.cuda.synchronize()
        nvtx.range_pop()
    print("ok", eng.last_info)


if __name__ == "__main__":
    main()

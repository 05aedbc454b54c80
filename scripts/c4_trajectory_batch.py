"""Config C4 at full size (BASELINE.json configs[3]): a batch of N independent Protein-Transduction trajectories
(K_m -> 0.3, the closed-form variant of tests/golden/protein_transduction_T99.npz) sharing one theta.

    python scripts/c4_trajectory_batch.py [--n 65536] [--iters 3]

Checks (sizes the oracle finishes in seconds):
  * pooled statistics / theta against a vectorised NumPy restatement (all trajectories stacked along the time axis:
    the statistics are plain sums over (trajectory, ODE, time));
  * the swept states of a few sampled trajectories against oracle/vgm_oracle.state_sweep with the pooled theta.
Prints one JSON line with the timings.  TEST/BENCH SCRIPT: imports oracle/ as the checker only.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import sympy as sym
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import variational_gradient_matching_for_dynamical_systems_b200 as pkg      # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200 import engine  # noqa: E402
from oracle import vgm_oracle as vo                                           # noqa: E402


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--sample", type=int, default=4)
    a = ap.parse_args()
    g = np.load(os.path.join(ROOT, "tests", "golden", "protein_transduction_T99.npz"))
    lines, params, names = [str(s) for s in g["lines"]], [str(s) for s in g["params"]], [str(s) for s in g["names"]]
    sc = [[int(v) for v in str(c).split(",") if v != ""] for c in g["couplings"]]
    observed = [str(s) for s in g["observed"]]
    K1, T, N = len(names), g["X0"].shape[0], a.n
    hidden1 = [u for u in range(K1) if names[u] not in observed]
    D, A = g["D"], g["A"]
    rng = np.random.default_rng(7)
    scale = 1.0 + 0.05 * rng.standard_normal((N, 1, K1))
    Xn = g["X0"][None, :, :] * scale                                   # [N, T, K1]
    X_all = np.ascontiguousarray(Xn.transpose(1, 0, 2).reshape(T, N * K1))

    t0 = time.time()
    model = pkg.OdeModel(lines, names, params, n_copies=N)
    hidden = [n * K1 + u for n in range(N) for u in hidden1]
    eng = engine.VGMEngine(model, D, A, hidden=hidden, couplings=sc, theta_mode="proxy")
    eng.set_states(X_all)
    torch.cuda.synchronize()
    setup_s = time.time() - t0

    # ---- pooled statistics, vectorised NumPy restatement --------------------------------------------------
    lin = vo.Linearisation(vo.parse_ode_lines(lines), sym.symbols(names), sym.symbols(params), couplings=sc)
    S_big = Xn.reshape(N * T, K1)                                      # trajectories stacked along time
    DX_big = np.einsum("st,ntk->nsk", D, Xn).reshape(N * T, K1)
    m, Sc = np.zeros(lin.p), np.zeros((lin.p, lin.p))
    for k in range(K1):
        B, b = lin.eval_B_b(k, S_big)
        m += B.T.dot(DX_big[:, k] - b)
        Sc += B.T.dot(B)
    theta_ref = np.linalg.solve(Sc, m)

    th = eng.theta_step().cpu().numpy()
    stats = eng.stats.cpu().numpy()
    out = {"n_trajectories": N, "T": T, "states": N * K1, "hidden_states": len(hidden), "setup_s": round(setup_s, 2)}
    out["relerr_local_mean"] = relerr(stats[:lin.p], m)
    out["relerr_local_scaling"] = relerr(stats[lin.p:].reshape(lin.p, lin.p), Sc)
    out["relerr_theta"] = relerr(th, theta_ref)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    eng.state_sweep()
    info = eng.last_info
    Xg = eng.get_states()                                              # [T, N*K1]
    pick = rng.choice(N, size=min(a.sample, N), replace=False)
    worst = 0.0
    for n in pick:
        Xr, _ = vo.state_sweep(Xn[n], th, lin, D, hidden1)
        worst = max(worst, relerr(Xg[:, n * K1:(n + 1) * K1], Xr))
    out["relerr_states_sampled"] = worst
    out["sweep_iterations"] = int(getattr(info, "iterations", -1))

    # ---- timing: full iterations on the device-resident states ---------------------------------------------
    eng.set_states(X_all)
    eng.iteration()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(a.iters):
        eng.iteration()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / a.iters
    out["ms_per_iteration"] = round(ms, 3)
    out["state_time_updates_per_s"] = len(hidden) * T / (ms * 1e-3)
    out["trajectories_per_s"] = N / (ms * 1e-3)
    # per-kernel-family device time of one more iteration (events around every launch: not a throughput number)
    import ctypes as C
    lib = eng.lib
    lib.vgm_profile_enable(1)
    info = eng.iteration()
    torch.cuda.synchronize()
    fam = {}
    for fid, name in enumerate(("dgemm", "tcgemm", "vec_other", "assemble", "ir_update", "ir_dir")):
        fl, by, fms, fn = C.c_double(), C.c_double(), C.c_double(), C.c_longlong()
        lib.vgm_profile_collect(fid, C.byref(fl), C.byref(by), C.byref(fms), C.byref(fn))
        if fn.value:
            fam[name] = {"ms": round(fms.value, 3), "launches": fn.value, "TFLOPs": round(fl.value / max(fms.value, 1e-9) * 1e-9, 2),
                         "GBs": round(by.value / max(fms.value, 1e-9) * 1e-6, 1)}
    lib.vgm_profile_enable(0)
    out["kernel_families"] = fam
    out["solver_iterations"] = info["iterations"] if isinstance(info, dict) else None
    ok = max(out["relerr_local_mean"], out["relerr_local_scaling"], out["relerr_theta"], worst) < 1e-9
    out["parity_1e-9"] = bool(ok)
    print(json.dumps(out))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

"""Time the LIVE, unmodified reference (through oracle/reference_shim.py) on its own analytical path.

    python scripts/time_live_reference.py            # build container only: needs /root/reference

BASELINE.md section 4: the live reference cannot reach K >= 1000 (O(K^2) SymPy set-up, (K+p+1)-argument callables),
so it is timed at T = 1000 for K in {10, 20, 40} -- Lorenz96, every second state observed and clamped, the call
sequence of VGM_for_Lorenz96.py:368-370 with optimizer='analytical' -- and the per-hidden-state cost (flat in K:
12 dense T^3 GEMMs + one gesv per hidden state, proxies...py:227,232,239,262) is what an extrapolation to the
benchmark shape multiplies.  Writes profiles/cpu_live_reference.json.  TEST/BENCH INFRASTRUCTURE: this script is the
only place the reference itself is timed; bench.py quotes the file (the reference cannot travel to the GPU box).
"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd
import sympy as sym

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vgm_oracle as vo                                                               # noqa: E402
from oracle.reference_shim import LocallyLinearODEs, Symbols, load_reference, writable_frame     # noqa: E402


def time_one(ref, K, T, dt, n_iter):
    names = ["_x_%d" % (i + 1) for i in range(K)]
    state, param = sym.symbols(names), sym.symbols(["_alpha"])
    symbols = Symbols(state, param)
    f = tempfile.NamedTemporaryFile("w", suffix="_ODEs.txt", delete=False)
    f.write("\n".join(vo.lorenz96_ode_lines(K)) + "\n")
    f.close()
    t0 = time.perf_counter()
    odes, odes_sym = ref.import_odes.import_odes(symbols, f.name)
    os.unlink(f.name)
    sc, _ = ref.import_odes.find_state_couplings_in_odes(odes_sym, symbols)
    lin = LocallyLinearODEs()
    lin.ode_param.B, lin.ode_param.b = ref.rewrite.rewrite_odes_as_linear_combination_in_parameters(odes, state, param)
    observed = names[0::2]
    lin.state.R, lin.state.r = ref.rewrite.rewrite_odes_as_linear_combination_in_states(
        odes, state, param, [s for s in state if str(s) in observed], sc)
    setup_s = time.perf_counter() - t0
    t, Xtrue = vo.lorenz96_trajectory(K, T, dt, seed=0)
    X = Xtrue.copy()
    X[:, 1::2] += 0.5 * np.random.default_rng(1).standard_normal((T, K // 2))
    # the reference's own kernel_function is 3 T^2 scalar Python calls (9.5 s at T = 1000, one-time set-up): the
    # operators come from the oracle here, the timed part is the coordinate-ascent iteration
    D, A, *_ = vo.kernel_function(t, t[:2], "rbf", [1.25 * dt, 1.0])
    obs = pd.DataFrame(X[:, 0::2], index=t, columns=observed)
    times = []
    for it in range(n_iter):
        sp = writable_frame(X, t, names)
        t0 = time.perf_counter()
        th = ref.proxies_strict.proxy_for_ode_parameters(sp, lin, D, A, param, None, odes, None, None, 'analytical')
        t1 = time.perf_counter()
        out = ref.proxies_strict.proxy_for_ind_states(sp, th, odes, None, lin, D, A, None, None, obs, None, sc, 5, True,
                                                      None, 'analytical')
        t2 = time.perf_counter()
        times.append((t1 - t0, t2 - t1))
        X = np.array(out.values, dtype=float)
    th_s = float(np.mean([a for a, _ in times]))
    sw_s = float(np.mean([b for _, b in times]))
    n_h = K // 2
    return {"K": K, "T": T, "hidden": n_h, "iterations_timed": n_iter, "symbolic_setup_s": round(setup_s, 3),
            "theta_update_s": round(th_s, 4), "state_sweep_s": round(sw_s, 4),
            "ms_per_hidden_state": round(1e3 * (th_s + sw_s) / n_h, 2),
            "state_time_updates_per_s": n_h * T / (th_s + sw_s)}


def main():
    ref = load_reference()
    try:
        from threadpoolctl import threadpool_info
        blas = [{k: d.get(k) for k in ("internal_api", "num_threads", "version")} for d in threadpool_info()]
    except Exception:  # noqa: BLE001
        blas = None
    T, dt = 1000, 0.05
    rows = [time_one(ref, K, T, dt, n_iter=2) for K in (10, 20, 40)]
    per = float(np.mean([r["ms_per_hidden_state"] for r in rows]))
    K_h = 50000
    out = {"what": "LIVE unmodified reference (VGM_modules, through oracle/reference_shim.py: matplotlib/plotting stubbed), "
                   "optimizer='analytical', Lorenz96, every second state observed + clamped, rbf l = 1.25 dt",
           "where": "build container (no GPU); NOT the GPU box -- the Python reference cannot travel there",
           "cores": os.cpu_count(), "blas": blas, "rows": rows,
           "mean_ms_per_hidden_state": round(per, 2),
           "extrapolated": {"label": "EXTRAPOLATED linearly in the number of hidden states from K <= 40 (the live "
                                     "reference cannot be set up at K = 100 000: O(K^2) SymPy, 100 001-argument callables)",
                            "config": "Lorenz96 K=100000 (50000 hidden), T=1000",
                            "seconds_per_iteration": round(per * 1e-3 * K_h, 1),
                            "state_time_updates_per_s": K_h * T / (per * 1e-3 * K_h)}}
    path = os.path.join(ROOT, "profiles", "cpu_live_reference.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

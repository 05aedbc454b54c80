"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` from the LIVE reference.

Run in the build container (needs ``/root/reference``):

    python -m oracle.make_golden

Every array stored here is an output of the unmodified reference functions
(``GP_regression.kernel_function``, ``rewrite_odes_as_local_linear_combinations.*``,
``proxies_for_ode_parameters_and_states.proxy_for_*`` with ``optimizer='analytical'``),
executed through ``oracle/reference_shim.py``; inputs are seeded (``default_rng``) and stored
next to the outputs, so the GPU box (which has no ``/root/reference``) can replay them.
``variant`` in each file says whether the strict or the patched proxies module produced it
(SURVEY.md section 8c item 5).
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import pandas as pd
import sympy as sym

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import vgm_oracle as vo                                    # noqa: E402
from oracle.reference_shim import (LocallyLinearODEs, Symbols, load_reference,  # noqa: E402
                                   writable_frame)

GOLD = os.path.join(ROOT, "tests", "golden")


def _write_odes(lines):
    f = tempfile.NamedTemporaryFile("w", suffix="_ODEs.txt", delete=False)
    f.write("\n".join(lines) + "\n")
    f.close()
    return f.name


def run_model(ref, lines, state_names, param_names, t, X0, observed_names, kernel, kparam,
              variant, n_iter, theta_for_states=None, clamp=True, couplings_override=None):
    """Run n_iter iterations of {theta step; state sweep} on the live reference."""
    state = sym.symbols(state_names)
    param = sym.symbols(param_names)
    symbols = Symbols(state, param)
    path = _write_odes(lines)
    odes, odes_sym = ref.import_odes.import_odes(symbols, path)
    os.unlink(path)
    sc, _ = ref.import_odes.find_state_couplings_in_odes(odes_sym, symbols)
    if couplings_override is not None:
        sc = couplings_override(sc)
    D, A, C, Cobs, Cso = ref.GP_regression.kernel_function(np.asarray(t), np.asarray(t), kernel, kparam)
    lin = LocallyLinearODEs()
    lin.ode_param.B, lin.ode_param.b = ref.rewrite.rewrite_odes_as_linear_combination_in_parameters(
        odes, state, param)
    obs_syms = [s for s in state if str(s) in observed_names]
    lin.state.R, lin.state.r = ref.rewrite.rewrite_odes_as_linear_combination_in_states(
        odes, state, param, obs_syms, sc)
    prox = ref.proxies_strict if variant == "strict" else ref.proxies_patched
    cols = list(map(str, state))
    obs = pd.DataFrame(np.asarray(X0)[:, [cols.index(n) for n in observed_names]], index=t,
                       columns=list(observed_names))
    X = np.array(X0, dtype=float)
    thetas, states = [], []
    for _ in range(n_iter):
        sp = writable_frame(X, t, cols)
        th = prox.proxy_for_ode_parameters(sp, lin, D, A, param, None, odes, None, None, 'analytical')
        thetas.append(th.values.ravel().copy())
        sp = writable_frame(X, t, cols)
        th_in = th if theta_for_states is None else pd.DataFrame(np.asarray(theta_for_states, float),
                                                                   columns=['value'])
        out = prox.proxy_for_ind_states(sp, th_in, odes, None, lin, D, A, None, None, obs, None, sc,
                                        5, clamp, None, 'analytical')
        X = np.array(out.values, dtype=float)
        states.append(X.copy())
    return dict(D=D, A=A, couplings=np.array([",".join(map(str, c)) for c in sc]),
                thetas=np.array(thetas), states=np.array(states))


def main():
    ref = load_reference()
    os.makedirs(GOLD, exist_ok=True)
    rng = np.random.default_rng(0)

    # ---- G1/G2: kernel_function -----------------------------------------------------
    cases = {
        "rbf_T40": (np.arange(0, 4, 0.1), np.arange(0, 4, 0.25), "rbf", [0.125, 1.0]),
        "rbf_T120": (np.arange(120) * 0.05, np.arange(0, 6, 0.3), "rbf", [0.0625, 1.3]),
        "rq_T30": (np.arange(0, 6, 0.2), np.array([0., .1, .5, .6, .8, 1, 1.4, 1.6, 2]), "RQ", [0.7, 0.22, 1.5]),
        "rq_default_T50": (np.arange(0, 4, 0.08), np.array([0., .1, .5, .6, .8, 1, 1.4, 1.6, 2]), "RQ", [0.1, 2, 5]),
    }
    out = {}
    for name, (t, t2, kt, kp) in cases.items():
        state = np.random.get_state()
        D, A, C, Cobs, Cso = ref.GP_regression.kernel_function(t, t2, kt, kp)
        np.random.set_state(state)
        out[name + "__t"] = t
        out[name + "__t2"] = t2
        out[name + "__type"] = np.array(kt)
        out[name + "__param"] = np.array(kp, dtype=float)
        for k, v in dict(D=D, A=A, C=C, Cobs=Cobs, Cso=Cso).items():
            out[name + "__" + k] = v
        out[name + "__condC"] = np.array(np.linalg.cond(C))
    np.savez_compressed(os.path.join(GOLD, "kernel_function.npz"), **out)

    # ---- Lorenz96, strict variant ---------------------------------------------------
    for tag, K, T, dtg, ell, clamp in (("lorenz96_K10_T40", 10, 40, 0.1, 0.125, True),
                                       ("lorenz96_K12_T100", 12, 100, 0.05, 0.0625, True),
                                       ("lorenz96_K8_T30_allhidden", 8, 30, 0.1, 0.125, False)):
        t, Xtrue = vo.lorenz96_trajectory(K, T, dtg, seed=1)
        names = ["_x_%d" % (i + 1) for i in range(K)]
        observed = names[0::2]
        X0 = Xtrue + 0.3 * rng.standard_normal(Xtrue.shape)
        if clamp:
            X0[:, 1::2] = Xtrue[:, 1::2] + 0.5 * rng.standard_normal(Xtrue[:, 1::2].shape)
        res = run_model(ref, vo.lorenz96_ode_lines(K), names, ["_alpha"], t, X0, observed,
                        "rbf", [ell, 1.0], "strict", 3, clamp=clamp)
        np.savez_compressed(os.path.join(GOLD, tag + ".npz"), t=t, X0=X0, kernel_param=np.array([ell, 1.0]),
                            observed=np.array(observed), names=np.array(names), variant=np.array("strict"),
                            clamp=np.array(clamp), **res)

    # ---- Lotka-Volterra, patched variant, both states hidden --------------------------
    t = np.arange(0, 4, 0.08)
    th = [2., 1., 4., 1.]
    x = np.array([5., 3.])
    Xs = []
    h = 0.001
    f = lambda x: np.array([th[0] * x[0] - th[1] * x[0] * x[1], -th[2] * x[1] + th[3] * x[0] * x[1]])
    for i in range(len(t)):
        Xs.append(x.copy())
        for _ in range(80):
            k1 = f(x); k2 = f(x + .5 * h * k1); k3 = f(x + .5 * h * k2); k4 = f(x + h * k3)
            x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    Xlv = np.array(Xs) + 0.1 * rng.standard_normal((len(t), 2))
    lv_lines = ["_theta_1 * _x_1 - _theta_2 * _x_1 * _x_2", "-_theta_3 * _x_2 + _theta_4 * _x_1 * _x_2"]
    res = run_model(ref, lv_lines, ["_x_1", "_x_2"], ["_theta_1", "_theta_2", "_theta_3", "_theta_4"],
                    t, Xlv, ["_x_1", "_x_2"], "rbf", [0.1, 1.0], "patched", 3, clamp=False)
    np.savez_compressed(os.path.join(GOLD, "lotka_volterra_T50.npz"), t=t, X0=Xlv, kernel_param=np.array([0.1, 1.0]),
                        lines=np.array(lv_lines), names=np.array(["_x_1", "_x_2"]),
                        params=np.array(["_theta_1", "_theta_2", "_theta_3", "_theta_4"]),
                        observed=np.array(["_x_1", "_x_2"]), variant=np.array("patched"), clamp=np.array(False), **res)

    # ---- Lorenz attractor, patched variant, hidden = _y -------------------------------
    T = 150
    t = np.arange(T) * 0.02
    x = np.array([1., 1., 1.])
    p3 = [10., 28., 8. / 3.]
    f = lambda x: np.array([p3[0] * (x[1] - x[0]), p3[1] * x[0] - x[1] - x[0] * x[2], x[0] * x[1] - p3[2] * x[2]])
    Xs = []
    h = 0.002
    for i in range(T):
        Xs.append(x.copy())
        for _ in range(10):
            k1 = f(x); k2 = f(x + .5 * h * k1); k3 = f(x + .5 * h * k2); k4 = f(x + h * k3)
            x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    Xla = np.array(Xs)
    Xla[:, 1] += 0.5 * rng.standard_normal(T)
    la_lines = ["_sigma * (_y - _x)", "_rho * _x - _y - _x * _z", "_x * _y - _alpha * _z"]
    res = run_model(ref, la_lines, ["_x", "_y", "_z"], ["_sigma", "_rho", "_alpha"], t, Xla, ["_x", "_z"],
                    "rbf", [0.025, 1.0], "patched", 3, clamp=True)
    np.savez_compressed(os.path.join(GOLD, "lorenz_attractor_T150.npz"), t=t, X0=Xla, kernel_param=np.array([0.025, 1.0]),
                        lines=np.array(la_lines), names=np.array(["_x", "_y", "_z"]),
                        params=np.array(["_sigma", "_rho", "_alpha"]), observed=np.array(["_x", "_z"]),
                        variant=np.array("patched"), clamp=np.array(True), **res)

    # ---- Protein transduction (K_m -> 0.3), patched; dS and R_pp clamped, couplings [] --
    T = 99
    t = np.arange(0, 99, 1.)
    k = [0.07, 0.6, 0.05, 0.3, 0.017]
    f = lambda x: np.array([-k[0] * x[0] - k[1] * x[0] * x[2] + k[2] * x[3], k[0] * x[0],
                            -k[1] * x[0] * x[2] + k[2] * x[3] + k[4] * x[4] / (0.3 + x[4]),
                            k[1] * x[0] * x[2] - (k[2] + k[3]) * x[3],
                            k[3] * x[3] - k[4] * x[4] / (0.3 + x[4])])
    x = np.array([1., 0., 1., 0., 0.])
    Xs = []
    h = 0.01
    for i in range(T):
        Xs.append(x.copy())
        for _ in range(100):
            k1 = f(x); k2 = f(x + .5 * h * k1); k3 = f(x + .5 * h * k2); k4 = f(x + h * k3)
            x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    Xpt = np.array(Xs) * (1 + 0.02 * rng.standard_normal((T, 5)))
    pt_lines = ["- _k_1 * _S - _k_2 * _S * _R + _k_3 * _RS", "_k_1 * _S",
                "- _k_2 * _S * _R + _k_3 * _RS + _V * _R_pp /(0.3 + _R_pp)",
                "_k_2 * _S * _R -(_k_3 + _k_4) * _RS", "_k_4 * _RS - _V * _R_pp / (0.3 + _R_pp)"]
    names = ["_S", "_dS", "_R", "_RS", "_R_pp"]
    params = ["_k_1", "_k_2", "_k_3", "_k_4", "_V"]

    def ovr(sc):
        sc = [list(c) for c in sc]
        sc[1] = []          # _dS appears in no RHS / clamped
        sc[2] = []          # _R  observed, clamped
        sc[4] = []          # _R_pp nonlinear in its own ODEs, clamped
        return sc
    res = run_model(ref, pt_lines, names, params, t, Xpt, ["_dS", "_R", "_R_pp"], "rbf", [1.25, 1.0],
                    "patched", 3, clamp=True, couplings_override=ovr)
    np.savez_compressed(os.path.join(GOLD, "protein_transduction_T99.npz"), t=t, X0=Xpt,
                        kernel_param=np.array([1.25, 1.0]), lines=np.array(pt_lines), names=np.array(names),
                        params=np.array(params), observed=np.array(["_dS", "_R", "_R_pp"]),
                        variant=np.array("patched"), clamp=np.array(True), **res)

    # ---- fitting_state_observations --------------------------------------------------
    K, T = 6, 40
    t, Xtrue = vo.lorenz96_trajectory(K, T, 0.1, seed=3)
    t2 = t[::4]
    names = ["_x_%d" % (i + 1) for i in range(K)]
    obs_cols = [0, 2, 4]
    Y = Xtrue[::4][:, obs_cols] + 0.2 * rng.standard_normal((len(t2), 3))
    D, A, C, Cobs, Cso = ref.GP_regression.kernel_function(t, t2, "rbf", [0.5, 2.0])
    obs = pd.DataFrame(Y, index=t2, columns=[names[c] for c in obs_cols])
    mean, inv_cov = ref.GP_regression.fitting_state_observations(obs, sym.symbols(names), 5.0, t, C, Cobs, Cso)
    np.savez_compressed(os.path.join(GOLD, "fitting_state_observations.npz"), t=t, t2=t2, Y=Y,
                        obs_cols=np.array(obs_cols), names=np.array(names), SNR=np.array(5.0),
                        kernel_param=np.array([0.5, 2.0]), C=C, Cobs=Cobs, Cso=Cso,
                        mean=mean.values, inv_cov=inv_cov)
    for fn in sorted(os.listdir(GOLD)):
        print(fn, os.path.getsize(os.path.join(GOLD, fn)) // 1024, "KiB")


if __name__ == "__main__":
    main()

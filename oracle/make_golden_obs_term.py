"""TEST INFRASTRUCTURE ONLY -- golden vectors for the opt-in OBSERVATION TERM of the state update.

    python -m oracle.make_golden_obs_term

The reference keeps the statement commented out (proxies_for_ode_parameters_and_states.py:265-266):

    global_mean[:,u] = np.squeeze(np.linalg.solve(local_scaling + state_pred_inv_cov,
                                  local_mean + np.dot(state_pred_inv_cov, state_pred_mean.iloc[:,u])))

(the MATLAB original has it active, archive/VGM_functions/proxy_for_ind_states.m:92-93).  ``reference_shim`` offers
``proxies_obs_term``: the shipped source with that statement put in the place of the active solve at :262 and nothing
else changed.  This script runs two sweeps of the live reference in that variant on the inputs of
``tests/golden/lorenz96_K10_T40.npz`` with a well-conditioned inverse covariance Q = (C + 0.5 I)^-1 and the GP-fit-like
mean mu = X0 + noise, and stores inputs + outputs in ``tests/golden/lorenz96_K10_T40_obs_term.npz``.
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

from oracle import vgm_oracle as vo                                                  # noqa: E402
from oracle.reference_shim import LocallyLinearODEs, Symbols, load_reference, writable_frame  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ref = load_reference()
    assert ref.proxies_obs_term is not None, "the commented-out statement was not found in the reference source"
    g = np.load(os.path.join(GOLD, "lorenz96_K10_T40.npz"))
    K = len(g["names"])
    names = [str(s) for s in g["names"]]
    t = g["t"]
    state, param = sym.symbols(names), sym.symbols(["_alpha"])
    symbols = Symbols(state, param)
    f = tempfile.NamedTemporaryFile("w", suffix="_ODEs.txt", delete=False)
    f.write("\n".join(vo.lorenz96_ode_lines(K)) + "\n")
    f.close()
    odes, odes_sym = ref.import_odes.import_odes(symbols, f.name)
    os.unlink(f.name)
    sc, _ = ref.import_odes.find_state_couplings_in_odes(odes_sym, symbols)
    D, A = g["D"], g["A"]
    lin = LocallyLinearODEs()
    lin.ode_param.B, lin.ode_param.b = ref.rewrite.rewrite_odes_as_linear_combination_in_parameters(odes, state, param)
    observed = [str(s) for s in g["observed"]]
    lin.state.R, lin.state.r = ref.rewrite.rewrite_odes_as_linear_combination_in_states(
        odes, state, param, [s for s in state if str(s) in observed], sc)
    rng = np.random.default_rng(77)
    C = vo.kernel_matrices(t, t[:2], "rbf", [0.125, 1.0])[0]
    Q = np.linalg.inv(C + 0.5 * np.eye(len(t)))
    Q = 0.5 * (Q + Q.T)
    X0 = np.array(g["X0"], dtype=float)
    mu = X0 + 0.2 * rng.standard_normal(X0.shape)
    obs = pd.DataFrame(X0[:, [names.index(n) for n in observed]], index=t, columns=observed)
    mu_df = pd.DataFrame(mu, index=t, columns=names)
    X = X0.copy()
    states = []
    for _ in range(2):
        sp = writable_frame(X, t, names)
        out = ref.proxies_obs_term.proxy_for_ind_states(sp, pd.DataFrame([8.0], columns=['value']), odes, None, lin, D, A,
                                                        mu_df, Q, obs, None, sc, 5, True, None, 'analytical')
        X = np.array(out.values, dtype=float)
        states.append(X.copy())
    np.savez_compressed(os.path.join(GOLD, "lorenz96_K10_T40_obs_term.npz"), Q=Q, mu=mu, X0=X0, D=D, A=A, t=t,
                        names=np.array(names), observed=np.array(observed), states=np.array(states),
                        variant=np.array("obs_term: :262 replaced by the reference's own commented statement :265-266"))
    # cross-check the oracle restatement right away
    lin_o = vo.Linearisation(vo.parse_ode_lines(vo.lorenz96_ode_lines(K)), state, param)
    hidden = [u for u in range(K) if names[u] not in observed]
    Xo = X0.copy()
    for it in range(2):
        Xo, _ = vo.state_sweep(Xo, [8.0], lin_o, D, hidden, obs_inv_cov=Q, obs_mean=mu)
        err = np.abs(Xo - states[it]).max() / np.abs(states[it]).max()
        print("sweep %d: oracle vs live reference (obs_term variant): %.2e" % (it, err))
        assert err < 1e-11


if __name__ == "__main__":
    main()

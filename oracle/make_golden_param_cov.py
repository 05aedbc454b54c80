"""TEST INFRASTRUCTURE ONLY -- golden vectors for the parameter "covariance" sum_k B_k^T A B_k.

The reference accumulates this matrix as ``global_cov`` (proxies_for_ode_parameters_and_states.py:60,79)
and never returns it, so it cannot be read off a reference call.  This script replays exactly the
statements of the reference's loop (:62-79) on the callables ``B[k]`` that the LIVE reference's own
``rewrite_odes_as_linear_combination_in_parameters`` produced, on the inputs (X0, A) already stored in
``tests/golden/<model>.npz``, and writes ``tests/golden/param_cov.npz`` (one p x p matrix per model).

    python -m oracle.make_golden_param_cov        # needs /root/reference
"""
from __future__ import annotations

import os
import sys

import numpy as np
import sympy as sym

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import vgm_oracle as vo                                    # noqa: E402
from oracle.make_golden import _write_odes                             # noqa: E402
from oracle.reference_shim import Symbols, load_reference              # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MODELS = ["lorenz96_K10_T40", "lorenz96_K12_T100", "lorenz96_K8_T30_allhidden", "lotka_volterra_T50",
          "lorenz_attractor_T150", "protein_transduction_T99"]


def main():
    ref = load_reference()
    out = {}
    for tag in MODELS:
        g = np.load(os.path.join(GOLD, tag + ".npz"), allow_pickle=True)
        names = [str(s) for s in g["names"]]
        if "lines" in g.files:
            lines, params = [str(s) for s in g["lines"]], [str(s) for s in g["params"]]
        else:
            lines, params = vo.lorenz96_ode_lines(len(names)), ["_alpha"]
        state = list(sym.symbols(names))
        param = sym.symbols(params) if len(params) > 1 else [sym.Symbol(params[0])]
        path = _write_odes(lines)
        odes, _ = ref.import_odes.import_odes(Symbols(state, param), path)
        os.unlink(path)
        B, _b = ref.rewrite.rewrite_odes_as_linear_combination_in_parameters(odes, state, param)
        state_proxy, eps_cov = np.array(g["X0"], dtype=float), np.array(g["A"], dtype=float)
        # --- the reference's statements, proxies...py:60,62,66,79 ---
        global_cov = np.zeros((len(param), len(param)))
        st = np.append(state_proxy, np.ones((state_proxy.shape[0], 1)), axis=1).T
        for k in range(len(B)):
            Bk = np.squeeze(B[k](*st)).T
            global_cov += Bk.T.dot(eps_cov.dot(Bk))
        out[tag] = global_cov
        print(tag, global_cov.shape, float(np.abs(global_cov).max()))
    np.savez_compressed(os.path.join(GOLD, "param_cov.npz"), **out)


if __name__ == "__main__":
    main()

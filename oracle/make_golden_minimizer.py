"""TEST INFRASTRUCTURE ONLY -- golden vectors for the 'minimizer' PARAMETER update.

    python -m oracle.make_golden_minimizer

Runs the LIVE reference's ``proxy_for_ode_parameters(optimizer='minimizer')`` (proxies...py:114-128: SciPy minimize on
squared_loss / squared_loss_gradient, :379-392) and its loss / gradient functions directly, on

* Lotka-Volterra, T = 50 (inputs of tests/golden/lotka_volterra_T50.npz), constraints None / 'shrinkage' /
  'nonnegative';
* Protein Transduction AS SHIPPED -- six parameters, not linear in _K_m (Protein_Transduction_ODEs.txt:5) -- on the
  states of tests/golden/protein_transduction_T99.npz: the model that has NO closed-form parameter update.

Stores theta, the loss and gradient at the start point and at theta, in tests/golden/minimizer_params.npz.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import sympy as sym

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import vgm_oracle as vo                                                          # noqa: E402
from oracle.reference_shim import Symbols, load_reference, writable_frame                   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def run(ref, lines, names, pnames, t, X0, D, tag, out):
    state, param = sym.symbols(names), sym.symbols(pnames)
    symbols = Symbols(state, param)
    f = tempfile.NamedTemporaryFile("w", suffix="_ODEs.txt", delete=False)
    f.write("\n".join(lines) + "\n")
    f.close()
    odes, odes_sym = ref.import_odes.import_odes(symbols, f.name)
    os.unlink(f.name)
    grad_param, _ = ref.import_odes.gradient_of_odes(symbols, odes_sym)
    P = ref.proxies_strict
    x0 = 0.1 * np.ones(len(pnames))
    DX = D.dot(X0)
    out[tag + "__lines"] = np.array(lines)
    out[tag + "__names"] = np.array(names)
    out[tag + "__params"] = np.array(pnames)
    out[tag + "__X0"] = X0
    out[tag + "__D"] = D
    out[tag + "__t"] = t
    out[tag + "__loss_x0"] = np.array(P.squared_loss(x0, X0, DX, odes, grad_param, 0.0))
    out[tag + "__grad_x0"] = np.array(P.squared_loss_gradient(x0, X0, DX, odes, grad_param, 0.0))
    lo, go = vo.squared_loss(x0, X0, D, list(odes_sym), state, param)
    assert abs(lo - out[tag + "__loss_x0"]) <= 1e-12 * abs(lo), (lo, out[tag + "__loss_x0"])
    assert np.max(np.abs(go - out[tag + "__grad_x0"])) <= 1e-11 * np.max(np.abs(go))
    for cons in (None, "shrinkage", "nonnegative"):
        sp = writable_frame(X0, t, names)
        th = P.proxy_for_ode_parameters(sp, None, D, None, param, None, odes, grad_param, cons, 'minimizer')
        th = th.values.ravel().astype(float)
        key = tag + "__theta_" + str(cons)
        out[key] = th
        alpha = 0.3 if cons == "shrinkage" else 0.0
        out[tag + "__loss_" + str(cons)] = np.array(P.squared_loss(th, X0, DX, odes, grad_param, alpha))
        print(tag, cons, th, float(out[tag + "__loss_" + str(cons)]))


def main():
    ref = load_reference()
    out = {}
    g = np.load(os.path.join(GOLD, "lotka_volterra_T50.npz"))
    run(ref, [str(s) for s in g["lines"]], [str(s) for s in g["names"]], [str(s) for s in g["params"]], g["t"],
        np.array(g["X0"], dtype=float), g["D"], "lv", out)
    g = np.load(os.path.join(GOLD, "protein_transduction_T99.npz"))
    lines = [ln.strip() for ln in open("/root/reference/Protein_Transduction_ODEs.txt").read().splitlines() if ln.strip()]
    run(ref, lines, ["_S", "_dS", "_R", "_RS", "_R_pp"], ["_k_1", "_k_2", "_k_3", "_k_4", "_V", "_K_m"], g["t"],
        np.array(g["X0"], dtype=float), g["D"], "pt6", out)
    np.savez_compressed(os.path.join(GOLD, "minimizer_params.npz"), **out)


if __name__ == "__main__":
    main()

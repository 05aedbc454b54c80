"""TEST INFRASTRUCTURE ONLY -- golden vectors of ``GP_regression.kernel_function`` for the kernel types beyond
rbf / RQ (GP_regression.py:181-231), produced by the LIVE reference.

    python -m oracle.make_golden_kernel_types      # needs /root/reference

Stored per case: inputs, the five returned arrays (D, A, C, Cobs, Cso) and cond(C).  The reference does not
return dC / ddC; they are pinned through D = dC C^-1 and A = ddC - D dC^T on cases whose C is well conditioned.
periodic / locally_periodic / exp_sin_squared differentiate sqrt((t0-t1)^2): the reference's derivative grids are
NaN on the diagonal, hence D and A are NaN -- the NaN pattern is stored as is.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.reference_shim import load_reference              # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    "matern05_T24": (np.arange(24) * 0.25, np.arange(0, 6, 0.7), "matern", [0.3, 0.5, 1.2]),
    "matern15_T24": (np.arange(24) * 0.25, np.arange(0, 6, 0.7), "matern", [0.35, 1.5, 0.9]),
    "matern25_T24": (np.arange(24) * 0.25, np.arange(0, 6, 0.7), "matern", [0.4, 2.5, 1.0]),
    "rbf_lin_T20": (np.arange(20) * 0.3, np.arange(0, 6, 0.9), "rbf+lin", [0.0, 0.35, 0.5, 0.2, 1.0, 1.1]),
    "sigmoid_T8": (0.5 + np.arange(8) * 0.9, np.array([0.7, 2.0, 4.4]), "sigmoid", [0.0, 0.3, 1.7, 1.0]),
    "sigmoid2_T8": (0.5 + np.arange(8) * 0.4, np.array([0.7, 2.0, 3.1]), "sigmoid2", [0.2, -0.4, 1.3]),
    "periodic_T16": (np.arange(16) * 0.21, np.array([0.1, 0.5, 1.9]), "periodic", [0.8, 1.7, 1.1]),
    "locally_periodic_T16": (np.arange(16) * 0.21, np.array([0.1, 0.5, 1.9]), "locally_periodic", [0.0, 0.9, 2.3, 1.2]),
    "exp_sin_squared_T16": (np.arange(16) * 0.21, np.array([0.1, 0.5, 1.9]), "exp_sin_squared", [0.7, 2.9, 0.8]),
}


def main():
    ref = load_reference()
    out = {}
    for name, (t, t2, kt, kp) in CASES.items():
        state = np.random.get_state()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                D, A, C, Cobs, Cso = ref.GP_regression.kernel_function(t, t2, kt, kp)
            except Exception as e:                      # noqa: BLE001
                print(name, "reference raised", type(e).__name__, e)
                continue
        np.random.set_state(state)
        out[name + "__t"], out[name + "__t2"] = t, t2
        out[name + "__type"], out[name + "__param"] = np.array(kt), np.array(kp, dtype=float)
        for k, v in dict(D=D, A=A, C=C, Cobs=Cobs, Cso=Cso).items():
            out[name + "__" + k] = v
        cond = np.linalg.cond(C)
        out[name + "__condC"] = np.array(cond)
        print("%-22s cond(C) %.2e  nan(D) %d/%d  |D|max %.3g" % (name, cond, int(np.isnan(D).sum()), D.size,
                                                                np.nanmax(np.abs(D)) if not np.isnan(D).all() else np.nan))
    np.savez_compressed(os.path.join(GOLD, "kernel_function_types.npz"), **out)


if __name__ == "__main__":
    main()

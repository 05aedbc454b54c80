"""NumPy walk-through of csrc/dense_solve.cu::factor_invert_diag_block: division-free (Bareiss-type) elimination of an 8 x 8
SPD block with power-of-two scaling one step behind, the row operations applied to an identity alongside, and the
inverse Cholesky factor recovered from one rsqrt per row at the end.  Lane r = row r; all() lanes run the same code."""
import math

import numpy as np


def pow2r(x):
    m, e = math.frexp(x)        # x = m 2^e, m in [0.5, 1)  ->  exponent of the [1, 2) form is e - 1
    return math.ldexp(1.0, -(e - 1))


def factor_invert(A):
    n = 8
    w = np.array(A, dtype=np.float64).copy()       # w[r, j]
    au = np.zeros((n, n))
    c = 1.0
    my_c = np.ones(n)
    piv = np.zeros(n)
    Rt = 1.0
    for k in range(n - 1):
        pkk = w[k, k]
        if k == 0:
            Rt = pow2r(pkk)
        piv[k] = pkk
        pkj = w[:, k].copy()                        # lane j's w[k]
        akj = au[k, :].copy()                       # lane k's au[j]
        for r in range(n):
            act = r > k
            pe, re, wk = (pkk, Rt, w[r, k] * Rt) if act else (1.0, 1.0, 0.0)
            for j in range(k + 1, n):
                w[r, j] = pe * (w[r, j] * re) - wk * pkj[j]
            for j in range(k):
                au[r, j] = pe * (au[r, j] * re) - wk * akj[j]
            au[r, k] = -wk * c
        c = c * (pkk * Rt)
        my_c[k + 1] = c
        Rt = pow2r(pkk)
    piv[n - 1] = w[n - 1, n - 1]
    inv = np.zeros((n, n))
    for r in range(n):
        scale = 1.0 / math.sqrt(my_c[r] * piv[r])
        for j in range(n):
            inv[r, j] = au[r, j] * scale if j < r else (my_c[r] * scale if j == r else 0.0)
    return inv, w


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    worst = 0.0
    for trial in range(200):
        B = rng.standard_normal((8, 8))
        s = 10.0 ** rng.uniform(-12, 12)
        A = (B @ B.T + 8 * rng.uniform(0.01, 10) * np.eye(8)) * s
        if trial % 3 == 0:                          # strongly graded diagonal
            d = 10.0 ** rng.uniform(-3, 3, size=8)
            A = A * d[:, None] * d[None, :]
        Aup = np.tril(A) + np.triu(rng.standard_normal((8, 8)), 1)   # the upper triangle is never to be trusted
        inv, w = factor_invert(Aup)
        ref = np.linalg.inv(np.linalg.cholesky(A))
        err = np.max(np.abs(inv - ref)) / np.max(np.abs(ref))
        res = np.max(np.abs(inv @ A @ inv.T - np.eye(8)))
        worst = max(worst, res)
        assert np.all(np.isfinite(w)), trial
        assert res < 1e-9 * max(1.0, np.linalg.cond(A) * 1e-4), (trial, err, res, np.linalg.cond(A))
    print("worst |inv A inv^T - I|:", worst)

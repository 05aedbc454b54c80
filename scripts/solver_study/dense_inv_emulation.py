"""NumPy walk-through of csrc/dense_solve.cu::dense_spd_solve_inv_kernel (block indices, work lists, fragment maths) --
the index logic of the kernel checked on the CPU before it goes to the GPU.  Usage: python dense_inv_emulation.py [T] [NW]"""
import math
import sys

import numpy as np


def blk_off(I, J):
    return I * (I + 1) // 2 + J


def run(T, NW, seed=0):
    rng = np.random.default_rng(seed)
    A0 = rng.standard_normal((T, T))
    A = A0 @ A0.T + T * np.eye(T)
    rhs = rng.standard_normal(T)
    nb = (T + 7) // 8
    n_tri = nb * (nb + 1) // 2
    W = np.zeros((n_tri + nb, 8, 8))
    # formation, warp by warp with the incremental (I, J) stepping of the kernel
    for warp in range(NW):
        I, J = 0, warp
        while J > I:
            J -= I + 1
            I += 1
        for b in range(warp, n_tri + nb, NW):
            if b >= n_tri:
                for k in range(8):
                    t = 8 * (b - n_tri) + k
                    W[b, 0, k] = rhs[t] if t < T else 0.0
            else:
                assert blk_off(I, J) == b
                for g in range(8):
                    for c in range(8):
                        i, j = 8 * I + g, 8 * J + c
                        if i < T:
                            W[b, g, c] = A[i, j] if j <= i else 0.0
                        else:
                            W[b, g, c] = 1.0 if i == j else 0.0
            J += NW
            while J > I:
                J -= I + 1
                I += 1

    def factor_invert(b):
        L = np.linalg.cholesky(np.tril(W[b]) + np.tril(W[b], -1).T)
        W[b] = np.linalg.inv(L)

    factor_invert(0)
    for pnl in range(nb):
        inv = W[blk_off(pnl, pnl)]
        for warp in range(NW):
            I = pnl + 1 + warp
            while I <= nb:
                W[blk_off(I, pnl)] = W[blk_off(I, pnl)] @ inv.T
                I += NW
        m = nb - 1 - pnl
        if m == 0:
            break
        done = set()
        W[blk_off(pnl + 1, pnl + 1)] -= W[blk_off(pnl + 1, pnl)] @ W[blk_off(pnl + 1, pnl)].T
        done.add((pnl + 1, pnl + 1))
        n_items = m * (m + 1) // 2 + m
        chunk = (n_items - 1 + NW - 2) // (NW - 1)
        for warp in range(1, NW):
            s0 = 1 + (warp - 1) * chunk
            cnt = min(chunk, n_items - s0)
            if cnt <= 0:
                continue
            i = int((math.sqrt(np.float32(8.0 * s0 + 1.0)) - 1.0) * 0.5)
            if i * (i + 1) // 2 > s0:
                i -= 1
            if (i + 1) * (i + 2) // 2 <= s0:
                i += 1
            I, J = pnl + 1 + i, pnl + 1 + (s0 - i * (i + 1) // 2)
            Ap, Cp, Bp = blk_off(I, pnl), blk_off(I, J), blk_off(J, pnl)
            jmax = min(I, nb - 1)
            while cnt > 0:
                assert Cp == blk_off(I, J) and Bp == blk_off(J, pnl) and Ap == blk_off(I, pnl), (pnl, I, J)
                assert (I, J) not in done
                done.add((I, J))
                W[Cp] -= W[Ap] @ W[Bp].T
                J += 1
                if J > jmax:
                    I += 1
                    J = pnl + 1
                    Ap = blk_off(I, pnl)
                    Cp = Ap + 1
                    Bp = blk_off(pnl + 1, pnl)
                    jmax = min(I, nb - 1)
                else:
                    Cp += 1
                    Bp += J
                cnt -= 1
        want = {(I, J) for I in range(pnl + 1, nb + 1) for J in range(pnl + 1, min(I, nb - 1) + 1)}
        assert done == want, (pnl, sorted(want - done), sorted(done - want))
        factor_invert(blk_off(pnl + 1, pnl + 1))
    # back substitution
    for pnl in range(nb - 1, -1, -1):
        acc = W[blk_off(nb, pnl)].copy()
        for I in range(nb - 1, pnl, -1):
            acc -= W[blk_off(nb, I)] @ W[blk_off(I, pnl)]
        W[blk_off(nb, pnl)] = acc @ W[blk_off(pnl, pnl)]
    x = np.array([W[blk_off(nb, t >> 3), 0, t & 7] for t in range(T)])
    ref = np.linalg.solve(A, rhs)
    return np.max(np.abs(x - ref)) / np.max(np.abs(ref))


if __name__ == "__main__":
    Ts = [int(sys.argv[1])] if len(sys.argv) > 1 else [2, 7, 8, 9, 33, 64, 99, 104, 150, 160]
    NWs = [int(sys.argv[2])] if len(sys.argv) > 2 else [2, 4, 8]
    for T in Ts:
        for NW in NWs:
            print(T, NW, run(T, NW))

"""Would a dot-free inner iteration pay?  Preconditioned Chebyshev (fixed coefficients from spectral bounds of S G S A)
against the PCG of ir4.py, same low-precision arithmetic (bf16-split operator, bf16 preconditioner, FP32 vectors).
A Chebyshev step needs no row dots, so on the GPU its vector updates could live in the epilogues of the two tcgen05
GEMMs (38 instead of 54 B per element and step, DESIGN.md section 8 i); the question is how many more steps it takes.
Run after study.py:   python cheb.py 56 20 64"""
import sys
import numpy as np
P = np.load("/tmp/vgm_study/prob.npz"); M = P["M"]; Lam = P["Lam"]; B = P["B"]; X0 = P["X0"]
n, T = B.shape; I = np.eye(T)
lbar = Lam.mean(); sh = 0.55 * lbar; c = 1.5 * lbar
G = np.linalg.inv(M + sh * I)
f32 = np.float32
def tile_band(A, band, BN=128, BK=64):
    out = np.zeros_like(A)
    for m0 in range(0, T, BN):
        lo = max(0, m0 - band) // BK * BK; hi = min(T, -(-min(T, m0 + BN + band) // BK) * BK)
        out[m0:m0 + BN, lo:hi] = A[m0:m0 + BN, lo:hi]
    return out
def bf16(x):
    x = np.asarray(x, dtype=np.float32); u = x.view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(np.float32)
def split(x):
    hi = bf16(x); return hi, bf16(np.asarray(x, dtype=np.float32) - hi)
bM, bC, bG = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
Mh, Ml = split(M); Gb = tile_band(bf16(G), bG)
Mh_main = tile_band(Mh, bM); Mh_corr = tile_band(Mh, bC); Ml_corr = tile_band(Ml, bC)
Lam32 = Lam.astype(f32); S = bf16(np.sqrt((sh + c) / (Lam + c)))
def mm(a, b): return a.astype(f32) @ b.astype(f32).T
def op_lp(v):
    vh, vl = split(v); return mm(vh, Mh_main) + mm(vl, Mh_corr) + mm(vh, Ml_corr) + Lam32 * (vh + vl)
def prec(r): return S * bf16(mm(bf16(S * r), Gb))

# spectral bounds of the preconditioned operator over ALL rows of the study problem (exact, FP64)
lo, hi = [], []
for i in range(n):
    s = np.sqrt((sh + c) / (Lam[i] + c))
    w = np.linalg.eigvals(((s[:, None] * G) * s[None, :]) @ (M + np.diag(Lam[i]))).real
    lo.append(w.min()); hi.append(w.max())
lo, hi = np.array(lo), np.array(hi)
print("spec(SGS A): min %.3f..%.3f  max %.3f..%.3f  kappa mean %.2f max %.2f" % (lo.min(), lo.max(), hi.min(), hi.max(), (hi / lo).mean(), (hi / lo).max()))

def inner_pcg(r0, nit):
    r = r0.astype(f32).copy(); d = np.zeros_like(r)
    z = prec(r); rz = np.sum(r * z, 1, dtype=np.float64); p = z.copy()
    for _ in range(nit):
        q = op_lp(p); al = (rz / np.sum(p * q, 1, dtype=np.float64)).astype(f32)[:, None]
        d += al * p; r = r - al * q
        z = prec(r); rzn = np.sum(r * z, 1, dtype=np.float64)
        p = z + (rzn / rz).astype(f32)[:, None] * p; rz = rzn
    return d
def inner_cheb(r0, nit, lmin, lmax):
    theta, delta = f32(0.5 * (lmax + lmin)), f32(0.5 * (lmax - lmin)); sig = theta / delta
    r = r0.astype(f32).copy(); x = np.zeros_like(r)
    rho = f32(1.0) / sig; d = prec(r) / theta
    for k in range(nit):
        x += d; r = r - op_lp(d)
        if k == nit - 1: break
        rho_n = f32(1.0) / (f32(2.0) * sig - rho)
        d = rho_n * rho * d + (f32(2.0) * rho_n / delta) * prec(r); rho = rho_n
    return x
def run(name, inner, maxout=14):
    x = X0.copy(); bn = np.linalg.norm(B, axis=1); hist = []
    for o in range(maxout):
        r = B - (x @ M.T + Lam * x); rn = np.linalg.norm(r, axis=1) / bn; hist.append(rn.max())
        if rn.max() <= 1e-13: break
        sc = np.linalg.norm(r, axis=1, keepdims=True)
        x = x + sc * inner(r / sc).astype(np.float64)
    print("%-34s corrections %2d   max relres per correction: %s" % (name, len(hist) - 1, " ".join("%.0e" % h for h in hist)))
    return len(hist) - 1
run("PCG, 10 steps per correction", lambda r: inner_pcg(r, 10))
for margin in (1.0, 1.1):
    lmin, lmax = lo.min() / margin, hi.max() * margin
    for nit in (10, 12, 14, 16):
        run("Chebyshev %2d steps, bounds x%.1f" % (nit, margin), lambda r: inner_cheb(r, nit, lmin, lmax))

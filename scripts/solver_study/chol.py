import numpy as np, sys, scipy.linalg as sl
P=np.load("/tmp/vgm_study/prob.npz"); M=P["M"];Lam=P["Lam"];B=P["B"];X0=P["X0"]; Xex=np.load("/tmp/vgm_study/xex.npy")
n,T=B.shape
idx=np.arange(T); dist=np.abs(idx[:,None]-idx[None,:])
b=int(sys.argv[1])
Mb=(M*(dist<=b)).astype(np.float32)
def run(maxout=8):
    x=X0.copy(); bn=np.linalg.norm(B,axis=1)
    # fp32 banded cholesky per row (use dense fp32 cholesky of banded matrix: same arithmetic up to ordering)
    Ls=[]
    for i in range(n):
        A32=(Mb+np.diag(Lam[i].astype(np.float32))).astype(np.float32)
        Ls.append(np.linalg.cholesky(A32))   # float32 LAPACK
    for o in range(maxout):
        r=B-(x@M.T+Lam*x); rn=np.linalg.norm(r,axis=1)/bn
        print("  outer %d: max relres %.2e median %.2e"%(o,rn.max(),np.median(rn)))
        if rn.max()<=1e-13: break
        for i in range(n):
            sc=np.linalg.norm(r[i]); rr=(r[i]/sc).astype(np.float32)
            y=sl.solve_triangular(Ls[i],rr,lower=True).astype(np.float32)
            d=sl.solve_triangular(Ls[i].T,y,lower=False).astype(np.float32)
            x[i]+=sc*d.astype(np.float64)
print("band",b); run()

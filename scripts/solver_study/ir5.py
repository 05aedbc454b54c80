"""Round-2 experiment: CG with RESIDUAL REPLACEMENT instead of restarted corrections.
Same arithmetic emulation as ir4.py (bf16-split operator, bf16 preconditioner with diagonal scaling, FP32 vectors), but
after every `nin` steps only the residual is replaced by the true FP64 residual (and x += sigma d is flushed to FP64):
the search direction p and r.z are KEPT, so the Krylov history survives.  sigma stays the first residual norm.
usage: python ir5.py bM bC bG nin [nin ...]   (after study.py)"""
import numpy as np, sys, os
P=np.load("/tmp/vgm_study/prob.npz"); M=P["M"];Lam=P["Lam"];B=P["B"];X0=P["X0"]
n,T=B.shape; I=np.eye(T)
lbar=Lam.mean(); sh=0.55*lbar; c=1.5*lbar
G=np.linalg.inv(M+sh*I)
def tile_band(A, band, BN=128, BK=64):
    out=np.zeros_like(A)
    for m0 in range(0,T,BN):
        lo=max(0,m0-band)//BK*BK; hi=min(T,-(-min(T,m0+BN+band)//BK)*BK)
        out[m0:m0+BN, lo:hi]=A[m0:m0+BN, lo:hi]
    return out
bM,bC,bG=int(sys.argv[1]),int(sys.argv[2]),int(sys.argv[3])
def bf16(x):
    x=np.asarray(x,dtype=np.float32); u=x.view(np.uint32).astype(np.uint64)
    r=((u+0x7FFF+((u>>16)&1))&0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)
def split(x):
    hi=bf16(x); lo=bf16(np.asarray(x,dtype=np.float32)-hi); return hi,lo
f32=np.float32
Mh,Ml=split(M); Gb=tile_band(bf16(G),bG)
Mh_main=tile_band(Mh,bM); Mh_corr=tile_band(Mh,bC); Ml_corr=tile_band(Ml,bC)
Lam32=Lam.astype(f32)
S=bf16(np.sqrt((sh+c)/(Lam+c)))
def mm(a,b): return (a.astype(f32)@b.astype(f32).T)
def op_lp(ph,pl): return mm(ph,Mh_main)+mm(pl,Mh_corr)+mm(ph,Ml_corr)+Lam32*(ph+pl)
def precond(r):
    rs=bf16(S*r); acc=bf16(mm(rs,Gb)); return S*acc, np.sum(rs*acc,1,dtype=np.float64)
def run(nin, replace, maxsteps=80):
    x=X0.copy(); bn=np.linalg.norm(B,axis=1)
    r64=B-(x@M.T+Lam*x)
    sig=np.linalg.norm(r64,axis=1,keepdims=True)
    r=(r64/sig).astype(f32); d=np.zeros_like(r)
    z,rz=precond(r); ph,pl=split(z); p=ph+pl
    steps=0
    while steps<maxsteps:
        for it in range(nin):
            q=op_lp(ph,pl)
            pq=np.sum(p*q,1,dtype=np.float64)
            al=(rz/pq).astype(f32)[:,None]
            d+=al*p; r=r-al*q; steps+=1
            if it==nin-1: break
            z,rzn=precond(r)
            be=(rzn/rz).astype(f32)[:,None]
            p=z+be*p; ph,pl=split(p); p=ph+pl; rz=rzn
        # flush to FP64 and look at the true residual
        x=x+sig*d.astype(np.float64); d[:]=0
        r64=B-(x@M.T+Lam*x)
        rn=np.linalg.norm(r64,axis=1)/bn
        print("  after %2d steps: max relres %.2e  median %.2e   rows above 1e-13: %d"%(steps,rn.max(),np.median(rn),(rn>1e-13).sum()))
        if rn.max()<=1e-13: break
        if replace:
            r=(r64/sig).astype(f32)                       # residual replacement: p and r.z survive
            z,rzn=precond(r)
            be=(rzn/rz).astype(f32)[:,None]
            p=z+be*p; ph,pl=split(p); p=ph+pl; rz=rzn
        else:
            sig=np.linalg.norm(r64,axis=1,keepdims=True)   # restart (what the solver does today)
            r=(r64/sig).astype(f32)
            z,rz=precond(r); ph,pl=split(z); p=ph+pl
for nin in [int(a) for a in sys.argv[4:]]:
    for replace in (0,1):
        print("bands",bM,bC,bG,"steps per pass",nin,"RESIDUAL REPLACEMENT" if replace else "restart"); run(nin,replace)

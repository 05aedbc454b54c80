import numpy as np, sys
P=np.load("/tmp/vgm_study/prob.npz"); M=P["M"];Lam=P["Lam"];B=P["B"];X0=P["X0"]; Xex=np.load("/tmp/vgm_study/xex.npy")
n,T=B.shape; I=np.eye(T)
lbar=Lam.mean(); sh=0.55*lbar; c=1.5*lbar
G=np.linalg.inv(M+sh*I)
idx=np.arange(T); dist=np.abs(idx[:,None]-idx[None,:])
def tile_band(A, band, BN=128, BK=64):
    # emulate the tile-granular k-range: column tile m0 uses k-blocks covering [m0-band, m0+BN+band)
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
RZB=int(sys.argv[4])
import os
ZB16=os.environ.get('ZB16')=='1'   # experiment: store z' = G (s o r) as bf16 instead of FP32
def gz(rs):
    a=mm(rs,Gb)
    return bf16(a) if ZB16 else a
def inner_pcg(r0, nit):
    r=r0.astype(f32).copy(); d=np.zeros_like(r)
    rs=bf16(S*r); acc=gz(rs); z=S*acc
    rz=np.sum((rs*acc) if RZB else (r*z),1,dtype=np.float64)
    ph,pl=split(z); p=ph+pl
    for it in range(nit):
        q=op_lp(ph,pl)
        pq=np.sum(p*q,1,dtype=np.float64)
        al=(rz/pq).astype(f32)[:,None]
        d+=al*p; r=r-al*q
        rs=bf16(S*r); acc=gz(rs); z=S*acc
        rzn=np.sum((rs*acc) if RZB else (r*z),1,dtype=np.float64)
        be=(rzn/rz).astype(f32)[:,None]
        p=z+be*p; ph,pl=split(p); p=ph+pl; rz=rzn
    return d
def run(nin, maxout=12):
    x=X0.copy()
    bn=np.linalg.norm(B,axis=1); tot=0
    for o in range(maxout):
        r=B-(x@M.T+Lam*x)
        rn=np.linalg.norm(r,axis=1)/bn
        print("  outer %d: max relres %.2e  median %.2e (inner total %d)"%(o,rn.max(),np.median(rn),tot))
        if rn.max()<=1e-13: break
        sc=np.linalg.norm(r,axis=1,keepdims=True)
        d=inner_pcg(r/sc,nin)
        x=x+sc*d.astype(np.float64); tot+=nin
for nin in [int(a) for a in sys.argv[5:]]:
    print("bands",bM,bC,bG,"rzb",RZB,"inner",nin); run(nin)

import numpy as np, scipy.linalg as sl, itertools
P=np.load("/tmp/vgm_study/prob.npz"); M=P["M"];Lam=P["Lam"]
n,T=Lam.shape; I=np.eye(T); lbar=Lam.mean()
rows=[0,5,11,17,23]
def kappa(Pinv_fn):
    ks=[]
    for i in rows:
        A=M+np.diag(Lam[i]); Pi=Pinv_fn(Lam[i])
        Pi=0.5*(Pi+Pi.T)
        w=np.linalg.eigvals(Pi@A).real
        ks.append(w.max()/w.min())
    return np.mean(ks), np.max(ks)
def one(sh_f=0.55,c_f=1.5):
    sh=sh_f*lbar; c=c_f*lbar; G=np.linalg.inv(M+sh*I)
    def f(lam):
        s=np.sqrt((sh+c)/(lam+c)); return (s[:,None]*G)*s[None,:]
    return f
print("one shift", kappa(one()))
Gcache={}
def Ginv(sh):
    if sh not in Gcache: Gcache[sh]=np.linalg.inv(M+sh*I)
    return Gcache[sh]
def two(f1,f2,c1,c2,thr,width):
    s1=f1*lbar; s2=f2*lbar; G1=Ginv(s1); G2=Ginv(s2)
    def f(lam):
        # smooth partition of unity in log(lam)
        x=(np.log(lam+1e-9)-np.log(thr*lbar))/width
        w2=0.5*(1+np.tanh(x)); w1=1-w2
        a1=np.sqrt(w1)*np.sqrt((s1+c1*lbar)/(lam+c1*lbar)); a2=np.sqrt(w2)*np.sqrt((s2+c2*lbar)/(lam+c2*lbar))
        return (a1[:,None]*G1)*a1[None,:]+(a2[:,None]*G2)*a2[None,:]
    return f
best=[]
for f1,f2,thr,width in itertools.product([0.1,0.2,0.3],[1.0,1.5,2.5],[0.4,0.7,1.0],[0.5,1.0]):
    for c1,c2 in [(0.5,3.0),(1.5,1.5),(0.3,4.0)]:
        k=kappa(two(f1,f2,c1,c2,thr,width)); best.append((k[0],k[1],f1,f2,thr,width,c1,c2))
best.sort()
for b in best[:8]: print(b)

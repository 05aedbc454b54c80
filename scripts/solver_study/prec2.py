import numpy as np
P=np.load("/tmp/vgm_study/prob.npz"); M=P["M"];Lam=P["Lam"]
T=M.shape[0]; I=np.eye(T)
rows=[0,3,5,11,17,20]
def kap(Pm,A):
    w=np.linalg.eigvals(Pm@A).real; return w.max()/w.min()
for sh in [3,8,15,30,45]:
    G=np.linalg.inv(M+sh*I)
    line=[]
    for c in [60,80,100,130,170]:
        ks=[]
        for i in rows:
            s=np.sqrt((sh+c)/(Lam[i]+c)); ks.append(kap((s[:,None]*G)*s[None,:],M+np.diag(Lam[i])))
        line.append("%.1f/%.1f"%(np.mean(ks),max(ks)))
    print("shift",sh,line)

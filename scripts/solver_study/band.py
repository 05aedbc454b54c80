import numpy as np
P=np.load("/tmp/vgm_study/prob.npz"); M=P["M"];G=P["G"]
T=M.shape[0]
def prof(A,name):
    mx=np.abs(A).max()
    out=[]
    for k in [0,1,2,4,8,16,24,32,48,64,96,128,192,256,384,512,768,999]:
        # max abs on k-th off-diagonals (interior rows only and overall)
        v=max(np.abs(np.diagonal(A,k)).max(),np.abs(np.diagonal(A,-k)).max())
        out.append("%d:%.1e"%(k,v/mx))
    print(name,"max",mx," ".join(out))
prof(M,"M"); prof(G,"G")
t=np.arange(T)*0.05; d=t[:,None]-t[None,:]; ell=0.0625
C=np.exp(-0.5*(d/ell)**2); dC=-(d/ell**2)*C
D=np.linalg.solve(C.T,dC.T).T
prof(D,"D")
# interior vs boundary: rows near edges
for k in [32,64,128]:
    print(k, np.abs(np.triu(M,k)).max()/np.abs(M).max(), "interior", np.abs(np.triu(M,k))[200:800].max()/np.abs(M).max())

"""Set-up of the solver study: Lorenz96 (K states, every second hidden), T=1000, RBF l=1.25 dt -> M, G, Lambda, rhs.
Writes /tmp/vgm_study/prob.npz and the exact solutions; the other scripts in this directory read them.
Stand-alone NumPy (no product or oracle imports); `mkdir -p /tmp/vgm_study` first."""
import numpy as np, time, sys
np.set_printoptions(linewidth=200, precision=3)
T=1000; dt=0.05; ell=0.0625; K=int(sys.argv[1]) if len(sys.argv)>1 else 48
t=np.arange(T)*dt
d=t[:,None]-t[None,:]
C=np.exp(-0.5*(d/ell)**2); dC=-(d/ell**2)*C
D=np.linalg.solve(C.T,dC.T).T
I=np.eye(T)
M=(I+D).T@(I+D)
# lorenz96 traj
def rhs(x,a=8.0): return (np.roll(x,-1,-1)-np.roll(x,2,-1))*np.roll(x,1,-1)-x+a
rng=np.random.default_rng(0)
x=8+rng.standard_normal(K); h=0.01
def step(x):
    k1=rhs(x);k2=rhs(x+.5*h*k1);k3=rhs(x+.5*h*k2);k4=rhs(x+h*k3);return x+h/6*(k1+2*k2+2*k3+k4)
for _ in range(500): x=step(x)
X=np.empty((T,K))
for i in range(T):
    X[i]=x
    for _ in range(5): x=step(x)
S=X.copy(); hid=np.arange(1,K,2); S[:,hid]+=0.5*rng.standard_normal((T,len(hid)))
a=8.0
DX=D@S
Lam=[];B=[]
for u in hid:
    s=lambda k: S[:,(u+k)%K]
    dx=lambda k: DX[:,(u+k)%K]
    ruu=a+(s(1)-s(-2))*s(-1)
    R1=s(2)-s(-1); r1=a-s(1)
    Rm=s(-2); rm=a-s(-1)-s(-3)*s(-2)
    R2=-s(1); r2=a+s(3)*s(1)-s(2)
    b=(I+D).T@ruu - (R1*(r1-dx(1))+Rm*(rm-dx(-1))+R2*(r2-dx(2)))
    Lam.append(R1**2+Rm**2+R2**2); B.append(b)
Lam=np.array(Lam); B=np.array(B)   # [n][T]
n=len(hid)
print("Lam mean",Lam.mean(),"max",Lam.max(),"min",Lam.min())
shift=Lam.mean()
G=np.linalg.inv(M+shift*I)
np.savez("/tmp/vgm_study/prob.npz",M=M,G=G,Lam=Lam,B=B,shift=shift,X0=S[:,hid].T.copy())
Xex=np.array([np.linalg.solve(M+np.diag(Lam[i]),B[i]) for i in range(n)])
np.save("/tmp/vgm_study/xex.npy",Xex)
w=np.linalg.eigvalsh(M); print("eig M",w[0],w[-1])
for i in range(3):
    Ai=M+np.diag(Lam[i]); wa=np.linalg.eigvalsh(Ai); print("cond A",wa[-1]/wa[0])
    import scipy.linalg as sl
    wg=sl.eigh(Ai,M+shift*I,eigvals_only=True); print("spec GA",wg[0],wg[1],wg[2],wg[-2],wg[-1])

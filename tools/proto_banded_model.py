import numpy as np, sys, time
import os
sys.path.insert(0,os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tph_dense as T
from global_racetrajectory_optimization_b200 import synth as S
F_SCALE=2.0
W=64
def pivots(diag, off):
    N=len(diag); d=np.zeros(N); dl=np.zeros(N)
    prev=diag[(-W)%N]
    for s in range(-W+1,N):
        i=s%N; prev=diag[i]-off[(i-1)%N]**2/prev
        if s>=0: d[i]=prev
    nxt=diag[(N-1+W)%N]
    for s in range(N-2+W,-1,-1):
        i=s%N; nxt=diag[i]-off[i]**2/nxt
        if s<N: dl[i]=nxt
    return d,dl
def trisolve(d, off, r):
    """cyclic solve Tri m = r via periodic LDL^T with warm-up. r: [N] or [N,k]"""
    N=len(d); y=np.zeros_like(r); m=np.zeros_like(r)
    prev=np.zeros_like(r[0])
    for s in range(-W,N):
        i=s%N; prev=r[i]-off[(i-1)%N]/d[(i-1)%N]*prev
        if s>=0: y[i]=prev
    nxt=np.zeros_like(r[0])
    for s in range(N-1+W,-1,-1):
        i=s%N; nxt=(y[i]-off[i]*nxt)/d[i]
        if s<N: m[i]=nxt
    return m
class Model:
    def __init__(self, rt, nv, scaling, kappa_bound, w_veh, b=32, BZ=40):
        N=rt.shape[0]; self.N=N; p=rt[:,:2]
        h=np.ones(N)
        for i in range(N-1): h[i+1]=h[i]/scaling[i]
        self.h=h; hm=np.roll(h,1)
        self.diag=2*(hm+h); self.off=h.copy()
        self.d,self.dl=pivots(self.diag,self.off)
        self.p=p; self.nv=nv
        self.m=trisolve(self.d,self.off,6*self.D2(p))
        a1=(np.roll(p,-1,axis=0)-p)-(h**2)[:,None]*(2*self.m+np.roll(self.m,-1,axis=0))/6
        self.xp,self.yp=a1[:,0],a1[:,1]
        c=(self.xp**2+self.yp**2)**-1.5
        self.sy=c*self.xp*h**2; self.sx=c*self.yp*h**2
        self.kref=self.sy*self.m[:,1]-self.sx*self.m[:,0]
        self.ub=rt[:,2]-w_veh/2; self.lb=-(rt[:,3]-w_veh/2)
        self.f=F_SCALE*self.Et(self.kref)
        self.b=b; self.BZ=BZ
    def D2(self,v):
        h=self.h; hm=np.roll(h,1)
        if v.ndim==1: return (np.roll(v,-1)-v)/h-(v-np.roll(v,1))/hm
        return (np.roll(v,-1,axis=0)-v)/h[:,None]-(v-np.roll(v,1,axis=0))/hm[:,None]
    def Z(self,v): return trisolve(self.d,self.off,6*self.D2(v))
    def Zt(self,v): return 6*self.D2(trisolve(self.d,self.off,v))
    def E(self,a): return self.sy*self.Z(self.nv[:,1]*a)-self.sx*self.Z(self.nv[:,0]*a)
    def Et(self,v): return self.nv[:,1]*self.Zt(self.sy*v)-self.nv[:,0]*self.Zt(self.sx*v)
    def Eband(self):
        """E[m, m+o] for o in [-BZ,BZ] -> array [N, 2BZ+1]"""
        N=self.N; BZ=self.BZ; h=self.h; hm=np.roll(h,1)
        tii=1.0/(self.d+self.dl-self.diag)
        rho_p=-np.roll(self.off,1)/self.dl; rho_m=-self.off/self.d
        # Tinv[m, m+o] for o in [-BZ-1, BZ+1] using symmetry: Tinv[m,k]=Tinv[k,m]; column m going outward
        TB=np.zeros((N,2*BZ+3))
        for m in range(N):
            TB[m,BZ+1]=tii[m]
            v=tii[m]
            for o in range(1,BZ+2):
                v=v*rho_p[(m+o)%N]; TB[m,BZ+1+o]=v     # Tinv[m+o, m] = Tinv[m+o-1,m]*rho+_{m+o}
            v=tii[m]
            for o in range(1,BZ+2):
                v=v*rho_m[(m-o)%N]; TB[m,BZ+1-o]=v     # Tinv[m-o, m]
        EB=np.zeros((N,2*BZ+1))
        for m in range(N):
            for o in range(-BZ,BZ+1):
                i=(m+o)%N
                z=6*(TB[m,BZ+1+o-1]/hm[i]-TB[m,BZ+1+o]*(1/hm[i]+1/h[i])+TB[m,BZ+1+o+1]/h[i])
                EB[m,BZ+o]=z*(self.sy[m]*self.nv[i,1]-self.sx[m]*self.nv[i,0])
        return EB
    def Hband(self):
        """H[i, i+k], k=0..b"""
        N=self.N; b=self.b; BZ=self.BZ; EB=self.Eband(); self.EB=EB
        HB=np.zeros((N,b+1))
        for i in range(N):
            for k in range(b+1):
                # sum_m E[m,i]E[m,i+k]; m ranges i+k-BZ .. i+BZ
                s=0.0
                for m in range(i+k-BZ,i+BZ+1):
                    mm=m%N
                    s+=EB[mm,BZ+(i-m)]*EB[mm,BZ+(i+k-m)]
                HB[i,k]=s
        return HB
def dense_from_band(HB):
    N,b1=HB.shape; M=np.zeros((N,N))
    for i in range(N):
        for k in range(b1):
            j=(i+k)%N
            M[i,j]=HB[i,k]; M[j,i]=HB[i,k]
    return M
if __name__=="__main__":
    N=int(sys.argv[1]) if len(sys.argv)>1 else 200
    rt=S.make_track(1,N)
    path=np.vstack((rt[:,:2],rt[0,:2]))
    cx,cy,A,nv=T.calc_splines(path)
    scaling=np.array([-A[4*i+2,4*i+5] for i in range(N-1)]+[A[4*N-2,1]])
    qp=T.assemble_min_curv(rt,nv,A,0.12,2.0)
    md=Model(rt,nv,scaling,0.12,2.0)
    print("xp err",np.abs(md.xp-qp['x_prime']).max(),"kref err",np.abs(md.kref-qp['k_kappa_ref']).max(),"f err",np.abs(md.f-qp['f']).max()/np.abs(qp['f']).max())
    a=np.random.default_rng(0).standard_normal(N)
    print("E op err",np.abs(md.E(a)-qp['E_kappa']@a).max()/np.abs(qp['E_kappa']@a).max())
    t=time.time(); HB=md.Hband(); print("Hband time",time.time()-t)
    Hd=dense_from_band(HB)
    # compare band entries
    H=qp['H']
    err=max(abs(H[i,(i+k)%N]-HB[i,k]) for i in range(N) for k in range(33))
    print("H band entry err",err/np.abs(H).max(), "tail", max(abs(H[i,(i+k)%N]) for i in range(N) for k in range(33,N//2))/np.abs(H).max())
    np.save('/tmp/HB.npy',HB)

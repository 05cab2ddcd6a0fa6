import numpy as np, sys, time
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tools'))
from oracle import tph_dense as T
from global_racetrajectory_optimization_b200 import synth as S
from proto_banded_model import Model, dense_from_band
from scipy.linalg import solve_triangular

class BlockCyc:
    """cyclic banded (b<=32) SPD solve via block tridiagonal Cholesky with 32x32 blocks + 32-row separator"""
    def __init__(self, HB, N):
        self.N=N; self.NA=N-32; self.nb=(self.NA+31)//32; self.HB=HB
        nb=self.nb
        # padded index -> real node (or -1)
        self.map=np.full(32*nb+32,-1)
        self.map[:self.NA]=np.arange(self.NA)
        self.map[32*nb:]=np.arange(self.NA,N)
    def entry(self, pi, pj, D):
        i=self.map[pi]; j=self.map[pj]
        if i<0 or j<0: return 1.0 if pi==pj else 0.0
        N=self.N
        k=(j-i)%N
        if k<=32 and (k<N-k or 2*k==N): v=self.HB[i,k]
        elif N-k<=32: v=self.HB[j,N-k]
        else: v=0.0
        if i==j: v+=D[i]
        return v
    def tile(self, rb, cb, D):
        return np.array([[self.entry(32*rb+r,32*cb+c,D) for c in range(32)] for r in range(32)])
    def factor(self, D):
        nb=self.nb
        self.Linv=[];self.Tm=[None];self.F=[]
        Ssep=self.tile(nb,nb,D)
        Tprev=None;Fprev=None
        for I in range(nb):
            A=self.tile(I,I,D)
            if I>0: A=A-Tprev@Tprev.T
            L=np.linalg.cholesky(A)
            Li=solve_triangular(L,np.eye(32),lower=True)
            self.Linv.append(Li)
            Y=self.tile(nb,I,D)
            FW=Y-(Fprev@Tprev.T if I>0 else 0)
            F=FW@Li.T
            self.F.append(F)
            Ssep=Ssep-F@F.T
            if I<nb-1:
                B=self.tile(I+1,I,D)
                Tn=B@Li.T
                self.Tm.append(Tn); Tprev=Tn
            Fprev=F
        L=np.linalg.cholesky(Ssep)
        self.LinvS=solve_triangular(L,np.eye(32),lower=True)
    def solve(self, g):
        nb=self.nb
        gp=np.zeros(32*nb+32)
        real=self.map>=0
        gp[real]=g[self.map[real]]
        y=np.zeros_like(gp)
        gS=gp[32*nb:].copy()
        for I in range(nb):
            t=gp[32*I:32*I+32].copy()
            if I>0: t-=self.Tm[I]@y[32*(I-1):32*I]
            y[32*I:32*I+32]=self.Linv[I]@t
            gS-=self.F[I]@y[32*I:32*I+32]
        yS=self.LinvS@gS
        x=np.zeros_like(gp)
        xS=self.LinvS.T@yS
        x[32*nb:]=xS
        for I in range(nb-1,-1,-1):
            t=y[32*I:32*I+32]-self.F[I].T@xS
            if I<nb-1: t-=self.Tm[I+1].T@x[32*(I+1):32*(I+2)]
            x[32*I:32*I+32]=self.Linv[I].T@t
        out=np.zeros(self.N)
        out[self.map[real]]=x[real]
        return out
if __name__=="__main__":
    N=int(sys.argv[1])
    rt=S.make_track(1,N)
    path=np.vstack((rt[:,:2],rt[0,:2]))
    cx,cy,A,nv=T.calc_splines(path)
    scaling=np.array([-A[4*i+2,4*i+5] for i in range(N-1)]+[A[4*N-2,1]])
    alpha_ref,err_ref=T.opt_min_curv(rt,nv,A,0.12,2.0)
    md=Model(rt,nv,scaling,0.12,2.0)
    HB=md.Hband(); Hd=dense_from_band(HB)
    bc=BlockCyc(HB,N)
    rng=np.random.default_rng(0)
    for dscale in [1e-8,1e-3,1e3]:
        D=dscale*10**rng.uniform(-4,4,N)
        bc.factor(D)
        g=rng.standard_normal(N)
        x=bc.solve(g)
        M=Hd+np.diag(D)
        xr=np.linalg.solve(M,g)
        print("dscale",dscale,"rel err",np.abs(x-xr).max()/np.abs(xr).max(),"resid",np.abs(M@x-g).max()/np.abs(g).max(), "cond", np.linalg.cond(M))

import numpy as np, sys, time
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tools'))
from oracle import tph_dense as T
from global_racetrajectory_optimization_b200 import synth as S
from proto_banded_model import Model, dense_from_band

def ipm_box(H, f, lb, ub, tol=1e-9, maxit=60, verbose=False, solve=None):
    N=len(f)
    a=0.5*(lb+ub); su=ub-a; sl=a-lb
    # initial multipliers: mu0 based
    g=H@a+f
    lu=np.maximum(-g,0)+1e-2*np.abs(g).max()+1e-8; ll=np.maximum(g,0)+1e-2*np.abs(g).max()+1e-8
    # alternative: complementarity-balanced
    hist=[]
    for it in range(maxit):
        rd=H@a+f+lu-ll
        mu=(su@lu+sl@ll)/(2*N)
        hist.append((it,np.abs(rd).max(),mu))
        scale=max(1.0,np.abs(f).max())
        if np.abs(rd).max()<=tol*scale and mu<=tol*1e0*max(1.0,abs(f@a))/N*0+tol: break
        D=lu/su+ll/sl
        M=H+np.diag(D)
        Lc=np.linalg.cholesky(M)
        sol=lambda r: np.linalg.solve(Lc.T,np.linalg.solve(Lc,r))
        # predictor
        rhs=-rd-( -lu) + (-ll)   # sigma=0: -(rd) - (0 - su*lu)/su + (0 - sl*ll)/sl = -rd + lu - ll
        da=sol(rhs)
        dlu=(-su*lu+lu*da)/su; dll=(-sl*ll-ll*da)/sl
        dsu=-da; dsl=da
        def maxstep(v,dv):
            m=dv<0
            return min(1.0,(-v[m]/dv[m]).min()) if m.any() else 1.0
        ap=min(maxstep(su,dsu),maxstep(sl,dsl)); ad=min(maxstep(lu,dlu),maxstep(ll,dll))
        mu_aff=((su+ap*dsu)@(lu+ad*dlu)+(sl+ap*dsl)@(ll+ad*dll))/(2*N)
        sigma=(mu_aff/mu)**3
        cu=dsu*dlu; cl=dsl*dll
        tu=sigma*mu-su*lu-cu; tl=sigma*mu-sl*ll-cl
        rhs=-rd-tu/su+tl/sl
        da=sol(rhs)
        dlu=(tu+lu*da)/su; dll=(tl-ll*da)/sl
        dsu=-da; dsl=da
        eta=max(0.995,1-mu) if False else 0.995
        ap=min(1.0,eta*min(maxstep(su,dsu)/1.0,maxstep(sl,dsl)) if True else 1)
        ap=eta*min(maxstep(su,dsu),maxstep(sl,dsl)); ap=min(ap,1.0)
        ad=min(1.0,eta*min(maxstep(lu,dlu),maxstep(ll,dll)))
        a=a+ap*da; su=ub-a; sl=a-lb
        lu=lu+ad*dlu; ll=ll+ad*dll
        if verbose: print(it, "rd %.2e mu %.2e sigma %.2e ap %.3f ad %.3f"%(np.abs(rd).max(),mu,sigma,ap,ad))
    return a, lu, ll, it, hist

if __name__=="__main__":
    N=int(sys.argv[1]) if len(sys.argv)>1 else 200
    for seed in [1,2,3]:
        rt=S.make_track(seed,N)
        path=np.vstack((rt[:,:2],rt[0,:2]))
        cx,cy,A,nv=T.calc_splines(path)
        scaling=np.array([-A[4*i+2,4*i+5] for i in range(N-1)]+[A[4*N-2,1]])
        alpha_ref,err_ref=T.opt_min_curv(rt,nv,A,0.12,2.0)
        md=Model(rt,nv,scaling,0.12,2.0)
        HB=md.Hband(); Hd=dense_from_band(HB)
        for tol in [1e-6,1e-8,1e-10]:
            a,lu,ll,it,hist=ipm_box(Hd,md.f,md.lb,md.ub,tol=tol,verbose=(seed==1 and tol==1e-10))
            print(seed,"tol",tol,"iters",it,"alpha err",np.abs(a-alpha_ref).max()/np.abs(alpha_ref).max(), "kappa max", np.abs(md.kref+md.E(a)).max())

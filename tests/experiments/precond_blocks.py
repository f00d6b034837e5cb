"""Host experiment behind the choice of the block-Jacobi block size: PCG iterations on the Schur
complement of the oracle's Hessian for preconditioner blocks of 1..32 consecutive poses. Test infrastructure (uses the oracle)."""
import sys, time; import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
from oracle.cpu_oracle import Oracle, JAC_ANALYTIC
from sparse_gslam_b200 import graphgen as gg
import importlib.util
spec=importlib.util.spec_from_file_location("tgp",os.path.join(ROOT, 'tests', 'test_gpu_parity.py')); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m)

def schur(H, b, nP3, lam):
    n=H.shape[0]
    Hd = H + lam*sp.identity(n, format='csr')
    Hpp=Hd[:nP3,:nP3].tocsc(); Hpl=Hd[:nP3,nP3:].tocsc(); Hll=Hd[nP3:,nP3:].tocsc()
    if Hll.shape[0]==0: return Hpp.tocsr(), b[:nP3]
    # Hll block diagonal 2x2 -> invert via splu (cheap)
    Hll_inv = spl.inv(Hll)
    S = (Hpp - Hpl @ Hll_inv @ Hpl.T).tocsr()
    bt = b[:nP3] - Hpl @ (Hll_inv @ b[nP3:])
    return S, bt

def pcg(S, b, Mapply, tol=1e-10, maxit=20000):
    x=np.zeros_like(b); r=b.copy(); z=Mapply(r); p=z.copy(); rz=r@z; rz0=rz
    for it in range(maxit):
        if rz <= tol*tol*rz0: return it
        Ap=S@p; a=rz/(p@Ap); x+=a*p; r-=a*Ap; z=Mapply(r); rzn=r@z; p=z+(rzn/rz)*p; rz=rzn
    return maxit

def block_jacobi(S, bs):
    n=S.shape[0]; blocks=[]
    S=S.tocsr()
    starts=list(range(0,n,bs))
    facs=[]
    for s0 in starts:
        e=min(n,s0+bs)
        D=S[s0:e,s0:e].toarray()
        facs.append(np.linalg.inv(D))
    def apply(r):
        z=np.empty_like(r)
        for s0,F in zip(starts,facs):
            e=min(n,s0+bs); z[s0:e]=F@r[s0:e]
        return z
    return apply

for name,g in (("c1",gg.make("c1")),("c2",gg.make("c2")),("c3",gg.make("c3")),("c5s",gg.make_c5(rows=80,cols=80))):
    o=Oracle(g); o.initialize_optimization()
    st=o.structure(); lin=o.linearize(JAC_ANALYTIC)
    H=m._blocks_to_csr(st, lin["H"])
    nP3=3*int((st["kind"]==0).sum())
    md=abs(H.diagonal()).max()
    for lam in (1e-5*md, 1e-5*md/3**8, 1e-5*md/3**14):
        S,bt=schur(H, lin["b"], nP3, lam)
        res=[]
        for k in (1,2,4,8,16,32):
            t=time.time(); it=pcg(S,bt,block_jacobi(S,3*k)); res.append((k,it))
        print(name, "lam %.2e"%lam, res, flush=True)

# Probe (CPU, NumPy): backward error of a blocked solve whose diagonal blocks are replaced by their inverses rounded once to the element type,
# for block orders 128..2048, both input recipes (reference: test/unified_rectrxm.jl:20-27; scaled: SURVEY.md 8(d)).  python probes/inv_block_accuracy.py 4096
import numpy as np, sys
def make(n,m,recipe,seed=0):
    r=np.random.RandomState(seed)
    if recipe=="ref":
        A=np.tril(r.rand(n,n)+1)+10*np.eye(n); B=r.rand(n,m)+1
    else:
        A=np.tril((2*r.rand(n,n)-1)/np.sqrt(n),-1)+np.diag(1+r.rand(n)); B=r.rand(n,m)+1
    return A,B
def solve_blocked(A,B,IB,dt):
    n=A.shape[0]; Ah=A.astype(dt); X=B.astype(dt).copy()
    A64=Ah.astype(np.float64)
    for o in range(0,n,IB):
        inv=np.linalg.inv(A64[o:o+IB,o:o+IB]); inv=np.tril(inv).astype(dt)
        X[o:o+IB]=(inv.astype(np.float32)@X[o:o+IB].astype(np.float32)).astype(dt)
        if o+IB<n:
            # one blocked update of all trailing rows (left-looking vs recursion differ only in rounding order)
            X[o+IB:]=(X[o+IB:].astype(np.float32)-Ah[o+IB:,o:o+IB].astype(np.float32)@X[o:o+IB].astype(np.float32)).astype(dt)
    return Ah.astype(np.float64),X.astype(np.float64)
n=int(sys.argv[1]); m=32
for recipe in ("ref","scaled"):
    A,B=make(n,m,recipe)
    for dt in (np.float16,np.float32):
        for IB in (128,512,1024,2048):
            if IB>n: continue
            Ah,X=solve_blocked(A,B,IB,dt)
            B0=B.astype(dt).astype(np.float64)
            be=np.linalg.norm(Ah@X-B0)/(np.linalg.norm(Ah)*np.linalg.norm(X)+np.linalg.norm(B0))
            print(recipe,dt.__name__,IB,"berr %.2e"%be, "cond blk %.1f"%np.linalg.cond(Ah[:IB,:IB]))

"""CPU-only study (NumPy): error of V = K_nm L^-T on an ill-conditioned K_mm (486 inducing points on a line, SqExponential) for fp32 FMA, 3xTF32 with
round-toward-zero / round-to-nearest fp32 accumulation per 8- or 32-deep MMA step (a model of the tensor cores truncating accumulator),
a blocked triangular solve and chunked accumulation.  Basis of DESIGN section 3 "Precision policy" and of the next-step note in section 9.
    python tests/studies/tf32x3_accumulation_model.py > profiles/r2/studies/tf32x3_accumulation_model.txt"""
import numpy as np, scipy.linalg as sl
rng=np.random.default_rng(3)
def tf32(x):
    u=x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u=(u+0x1000)&0xFFFFE000
    return u.astype(np.uint32).view(np.float32)
def split(x):
    x=x.astype(np.float32); hi=tf32(x); lo=tf32((x-hi).astype(np.float32)); return hi.astype(np.float64),lo.astype(np.float64)
def rz32(x):  # fp64 -> fp32 round toward zero
    y=x.astype(np.float32)
    over=np.abs(y.astype(np.float64))>np.abs(x)
    y2=np.nextafter(y,np.float32(0))
    return np.where(over,y2,y)
def gemm3(A,Bm,mode,kc=8):
    # A [M,K], Bm [N,K]; C = A Bm^T with 3xTF32; accumulate per k8 step
    Ah,Al=split(A); Bh,Bl=split(Bm)
    M,K=A.shape; N=Bm.shape[0]
    acc=np.zeros((M,N),np.float32)
    for k in range(0,K,kc):
        s=slice(k,k+kc)
        for (a,b) in ((Al,Bh),(Ah,Bl),(Ah,Bh)):
            p=a[:,s]@b[:,s].T
            t=acc.astype(np.float64)+p
            acc=rz32(t) if mode=="rz" else t.astype(np.float32)
    return acc.astype(np.float64)
n,D,m,sc=1149,1,486,3.0
X=rng.standard_normal((n,D)); Z=X[rng.permutation(n)[:m]].copy()
Xs,Zs=X*sc,Z*sc
K=np.exp(-0.5*((Xs[:,None,:]-Zs[None,:,:])**2).sum(-1))
Kmm=np.exp(-0.5*((Zs[:,None,:]-Zs[None,:,:])**2).sum(-1))+1e-5*np.eye(m)
L=np.linalg.cholesky(Kmm); Li=sl.solve_triangular(L,np.eye(m),lower=True)
K32=K.astype(np.float32).astype(np.float64); Li32=Li.astype(np.float32).astype(np.float64)
V=K32@Li32.T   # exact product of the fp32 inputs
Vsimt=(K32.astype(np.float32)@Li32.astype(np.float32).T).astype(np.float64)
r=lambda a: np.linalg.norm(a-V)/np.linalg.norm(V)
print("amp",np.sqrt(np.abs(np.linalg.inv(Kmm)).sum(1).max()))
print("fp32 blas",r(Vsimt))
for mode in ("rz","rn"):
    for kc in (8,32):
        print("3xtf32",mode,kc,r(gemm3(K32,Li32,mode,kc)))
print("input rounding effect", np.linalg.norm(V-K@Li.T)/np.linalg.norm(V))

print("---- blocked triangular solve with M_j = L_jj^-1 [-L_j0 .. -L_j,j-1, I]")
def blocked(K32, L, nb, mode, kc=8):
    m=L.shape[0]; Buf=K32.copy()
    for j0 in range(0,m,nb):
        j1=min(m,j0+nb)
        Ljj_inv=sl.solve_triangular(L[j0:j1,j0:j1],np.eye(j1-j0),lower=True)
        Mj=np.concatenate([-Ljj_inv@L[j0:j1,:j0], Ljj_inv],axis=1)   # (nb, j1)
        Mj32=Mj.astype(np.float32).astype(np.float64)
        A=Buf[:,:j1]
        if mode=="f32":
            Vj=(A.astype(np.float32)@Mj32.astype(np.float32).T).astype(np.float64)
        else:
            Vj=gemm3(A,Mj32,mode,kc)
        Buf[:,j0:j1]=Vj.astype(np.float32).astype(np.float64)
    return Buf
Vex=K@Li.T
r2=lambda a: np.linalg.norm(a-Vex)/np.linalg.norm(Vex)
print("direct 3xtf32 rz vs exact", r2(gemm3(K32,Li32,"rz",8)), " fp32 blas", r2(Vsimt))
for nb in (128,64,32):
    print("nb",nb,"rz",r2(blocked(K32,L,nb,"rz")),"rn",r2(blocked(K32,L,nb,"rn")),"f32",r2(blocked(K32,L,nb,"f32")))

print("---- chunked: RZ accumulation inside a chunk of C k-values (fresh accumulator), RN fp32 sum of the chunk results")
def chunked(A,Bm,C):
    out=np.zeros((A.shape[0],Bm.shape[0]),np.float32)
    for k in range(0,A.shape[1],C):
        part=gemm3(A[:,k:k+C],Bm[:,k:k+C],"rz",8).astype(np.float32)
        out=(out.astype(np.float64)+part).astype(np.float32)
    return out.astype(np.float64)
for C in (32,64,128,256):
    print("chunk",C,r2(chunked(K32,Li32,C)))

import numpy as np, sys
sys.path.insert(0,'/root/repo')
from pagmo2_b200 import synth
D=100
mr,os_c,s=synth.cec2014_tables(1,D)
M=mr[:D*D].reshape(D,D)
rng=np.random.default_rng(0)
def slices(V, nsl, axis_scale):
    # V: [rows, K] doubles; per-row power-of-two scale so that |v| < 2^e; balanced base-256 digits, most significant first
    mx=np.abs(V).max(axis=1)
    e=np.where(mx>0, np.floor(np.log2(np.where(mx>0,mx,1)))+1, 0).astype(np.int64)   # |v| < 2^e
    S=8*nsl-1
    Y=np.rint(np.ldexp(V, (S-e)[:,None].astype(np.int64))).astype(np.int64)       # |Y| <= 2^S
    bias=sum(0x80<<(8*i) for i in range(nsl))
    Yb=Y+bias
    dig=[(((Yb>>(8*(nsl-1-i)))&0xff)^0x80).astype(np.int8).astype(np.int64) for i in range(nsl)]  # wrong: xor then sign
    dig=[((((Yb>>(8*(nsl-1-i)))&0xff).astype(np.int64))-128) for i in range(nsl)]
    # check reconstruction
    rec=sum(d<<(8*(nsl-1-i)) for i,d in enumerate(dig))
    assert (rec==Y).all(), np.abs(rec-Y).max()
    return dig,e,S
def ozaki(Mm, Yv, nA, nB, maxg):
    dB,eB,SB=slices(Mm,nB,None)     # rows of M (outputs), per-row scale
    dA,eA,SA=slices(Yv,nA,None)     # individuals
    n=Yv.shape[0]
    z=np.zeros((n,Mm.shape[0]))
    acc=[np.zeros((n,Mm.shape[0]),dtype=np.int64) for _ in range(maxg+1)]
    for i in range(nA):
        for j in range(nB):
            if i+j<=maxg:
                acc[i+j]+=dA[i]@dB[j].T
    # value = sum_g acc_g * 2^{8(nA-1-i)+8(nB-1-j)} / 2^{SA-eA} / 2^{SB-eB}; weight exponent for group g: 8(nA+nB-2-g)
    for g in range(maxg,-1,-1):
        z+=np.ldexp(acc[g].astype(np.float64), 8*(nA+nB-2-g))
        assert np.abs(acc[g]).max()<2**31
    z=np.ldexp(z, (eA[:,None]-SA)+(eB[None,:]-SB))
    return z
for name,Yv in (("uniform", (rng.uniform(-100,100,(2000,D))-os_c[:D])), ("near-opt", rng.normal(0,1,(2000,D))), ("tiny", rng.normal(0,1e-6,(500,D)))):
    exact=np.array([[float(np.sum(np.array(Yv[i],dtype=np.longdouble)*np.array(M[j],dtype=np.longdouble))) for j in range(D)] for i in range(min(300,Yv.shape[0]))])
    ref64=Yv[:exact.shape[0]]@M.T
    for nA,nB,maxg in ((6,6,5),(7,6,6),(7,7,6),(6,6,6),(7,7,7)):
        z=ozaki(M,Yv[:exact.shape[0]],nA,nB,maxg)
        scale=np.abs(exact).max(axis=1,keepdims=True)
        err=np.abs(z-exact)/scale
        f_ex=(exact**2*10**(6*np.arange(D)/(D-1))).sum(1); f_oz=(z**2*10**(6*np.arange(D)/(D-1))).sum(1)
        nprod=sum(1 for i in range(nA) for j in range(nB) if i+j<=maxg)
        print(name,nA,nB,maxg,"prods",nprod,"max err/rowmax %.2e"%err.max(), "f1 rel %.2e"%np.abs(f_oz/f_ex-1).max(), " fp64 matmul err %.2e"%(np.abs(ref64-exact)/scale).max())

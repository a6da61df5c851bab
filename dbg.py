import sys, os, ctypes
sys.path.insert(0,'.'); sys.path.insert(0,'tests'); sys.path.insert(0,'oracle')
import numpy as np, torch
import heatsim2_b200 as hs
from heatsim2_b200 import _cabi
import problems
np.set_printoptions(linewidth=200, precision=3)
lib=_cabi.lib()
prob=problems.uniform_slab(hs,shape=(1,16,16))
P_,S=hs.setup(*prob["setup_args"]); pa=P_.plan; pa.ensure_device()
print(pa.chunk)
fwd,bwd,G=pa.chunk_tabs[1]
M,P=pa.chunk[1]; L=16
lo,dg,hi=pa.line_rows[1]
A=np.diag(dg[0])+np.diag(lo[0,1:],-1)+np.diag(hi[0,:-1],1)
st=ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def emul(d):
    u=0
    uu=np.zeros(L); yf=np.zeros(P); yl=np.zeros(P); x=np.zeros(L)
    for p in range(P):
        prev=0.0; acc=0.0
        for rr in range(p*M,min(L,(p+1)*M)):
            prev=d[rr]*fwd[u,rr,0]-fwd[u,rr,1]*prev
            uu[rr]=prev; acc+=fwd[u,rr,2]*prev
        yf[p]=acc; yl[p]=prev
    Y=np.concatenate([yf,yl])
    for p in range(P):
        al=G[u,p,0]@Y; be=G[u,p,1]@Y
        xn=be
        for rr in range(min(L,(p+1)*M)-1,p*M-1,-1):
            xn=(uu[rr]-al*bwd[u,rr,0])-bwd[u,rr,1]*xn
            x[rr]=xn
    return x
for imp in (0,5,8,15):
    W0n=np.zeros(pa.shape); W0n[0,imp,:]=1.0
    wa=torch.from_numpy(W0n).cuda()
    _cabi.check(lib.hs2_sweep_y(pa._handle, wa.data_ptr(), st)); torch.cuda.synchronize()
    wa=wa.cpu().numpy()[0,:,3]
    xe=emul(W0n[0,:,3]); xr=np.linalg.solve(A,W0n[0,:,3])
    print("imp",imp,"gpu-exact",wa-xr)
    print("       emul-exact",xe-xr)

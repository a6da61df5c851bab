import sys
sys.path[:0]=['tests','oracle','.']
import numpy as np, torch
import heatsim2_b200 as hs, problems, adi_oracle, util
for name,kw in (("uniform_slab",dict(shape=(4,8,64))),("steelonfoam",dict(nz=24,ny=20,nx=48,nsteps=3)),("steelonwater",dict(nz=9,ny=14,nx=128)),("sources_demo",dict(nz=12,ny=10,nx=16))):
    prob=problems.ALL[name](hs,**kw)
    P,S=hs.setup(*prob["setup_args"]); O=adi_oracle.setup(*prob["setup_args"])
    T=np.array(prob["T0"])
    for it in range(3):
        t=prob["t0"]+it*prob["dt"]
        want=O.step(t,prob["dt"],T)
        got=hs.run_adi_steps(P,S,t,prob["dt"],torch.from_numpy(T).cuda(),prob["volumetric_elements"],prob["volumetric"]).cpu().numpy()
        print(name,kw,it,P.plan.last_kernels(),"relerr %.2e"%util.relerr(got,want),flush=True)
        T=want

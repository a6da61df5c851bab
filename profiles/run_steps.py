"""Minimal driver for ncu: N source-free ADI steps on the bench workload.
    ncu ... python profiles/run_steps.py [grid] [steps]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import heatsim2_b200 as hs
import problems

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prob = problems.uniform_slab(hs, shape=(grid, grid, grid), random_T0=False)
P, S = hs.setup(*prob["setup_args"])
g = torch.Generator(device="cuda").manual_seed(1234)
Ta = torch.rand(P.plan.shape, dtype=torch.float64, device="cuda", generator=g)
Tb = torch.empty_like(Ta)
for it in range(1, steps + 1):
    hs.run_adi_steps(P, S, it * prob["dt"], prob["dt"], Ta, prob["volumetric_elements"], prob["volumetric"], out=Tb)
    Ta, Tb = Tb, Ta
torch.cuda.synchronize()
print("done", float(Ta.mean()))

"""One x-sweep of the march kernel on a small problem (debugging aid, e.g. under compute-sanitizer):
    HS2_XM_R=4 compute-sanitizer --tool memcheck --kernel-name regex:sweep_xm python scripts/xm_one.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import heatsim2_b200 as hs
from heatsim2_b200 import _cabi
import problems

name = sys.argv[1] if len(sys.argv) > 1 else "steelonfoam"
prob = problems.ALL[name](hs)
lib = _cabi.lib()
T = torch.rand(tuple(np.array(prob["T0"]).shape), dtype=torch.float64, device="cuda")
out = {}
for flags in (0, 2):
    P, S = hs.setup(*prob["setup_args"])
    P.plan.flags = flags
    P.plan.ensure_device(T.device)
    W = torch.full_like(T, float("nan"))
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _cabi.check(lib.hs2_sweep_x(P.plan._handle, T.data_ptr(), W.data_ptr(), None, None, None, st))
    torch.cuda.synchronize()
    out[flags] = W
    print("flags", flags, P.plan.x_kernel, "chunk", P.plan.chunk[0], "finite", bool(torch.isfinite(W).all()), flush=True)
print("relerr", float((out[0] - out[2]).abs().max() / out[0].abs().max()))

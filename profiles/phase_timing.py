"""Phase timing of the x-sweep (needs a library built with -DHS2_PHASE_TIMING):
    python heatsim2_b200/build.py --force -DHS2_PHASE_TIMING && python profiles/phase_timing.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import heatsim2_b200 as hs
from heatsim2_b200 import _cabi
import problems
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prob = problems.uniform_slab(hs, shape=(grid, grid, grid), random_T0=False)
P, S = hs.setup(*prob["setup_args"])
plan = P.plan
plan.ensure_device()
lib = _cabi.lib()
Ta = torch.rand(plan.shape, dtype=torch.float64, device="cuda")
W = torch.empty_like(Ta)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
lib.hs2_debug_phase.argtypes = [ctypes.c_void_p, ctypes.c_int]
for it in range(3):
    _cabi.check(lib.hs2_sweep_x(plan._handle, Ta.data_ptr(), W.data_ptr(), None, None, None, st))
torch.cuda.synchronize()
lib.hs2_debug_phase(None, 1)
n = 5
for it in range(n):
    _cabi.check(lib.hs2_sweep_x(plan._handle, Ta.data_ptr(), W.data_ptr(), None, None, None, st))
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 16)()
lib.hs2_debug_phase(out, 0)
nblocks = grid * ((grid + 7) // 8) * n   # tiles
names = ["phase1 rhs", "sync", "load+fwd", "sync", "interface", "sync", "bwd+sts", "sync", "phase3 store"]
tot = sum(out[:9])
for i, nm in enumerate(names):
    print("%-14s %8.0f cycles/block  %5.1f%%" % (nm, out[i] / nblocks, 100.0 * out[i] / tot))
print("total %.0f cycles/block" % (tot / nblocks))

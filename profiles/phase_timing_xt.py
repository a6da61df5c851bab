"""Phase timing of the TMA-fed x sweep (kernels_xt.cu): cycles thread 0 of every block spends between
the marks, per patch.  Needs a library built with -DHS2_PHASE_TIMING:
    python -m heatsim2_b200.build --force -DHS2_PHASE_TIMING && python profiles/phase_timing_xt.py"""
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
lib.hs2_debug_phase_xt.argtypes = [ctypes.c_void_p, ctypes.c_int]
for it in range(3):
    _cabi.check(lib.hs2_sweep_x(plan._handle, Ta.data_ptr(), W.data_ptr(), None, None, None, st))
torch.cuda.synchronize()
lib.hs2_debug_phase_xt(None, 1)
n = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(n):
    _cabi.check(lib.hs2_sweep_x(plan._handle, Ta.data_ptr(), W.data_ptr(), None, None, None, st))
e1.record()
torch.cuda.synchronize()
print("kernel time with the marks compiled in: %.4f ms" % (e0.elapsed_time(e1) / n))
out = (ctypes.c_ulonglong * 16)()
lib.hs2_debug_phase_xt(out, 0)
ntiles = ((grid + 1) // 2) * ((grid + 3) // 4) * n
names = ["ids+lid issue", "wait C", "phase1 rhs (+wait Z)", "sync1", "issue loads+forward", "sync2", "interface", "sync3",
         "backward", "store"]
tot = sum(out[:10])
print(plan.last_kernels())
for i, nm in enumerate(names):
    print("%-24s %8.0f cycles/patch  %5.1f%%" % (nm, out[i] / ntiles, 100.0 * out[i] / tot))
print("total %.0f cycles/patch/block" % (tot / ntiles))

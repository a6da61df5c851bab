"""Phase timing of the persistent strided kernels (library built with -DHS2_PHASE_TIMING)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import heatsim2_b200 as hs
from heatsim2_b200 import _cabi
import problems
grid = 512
prob = problems.uniform_slab(hs, shape=(grid, grid, grid), random_T0=False)
P, S = hs.setup(*prob["setup_args"])
plan = P.plan
plan.ensure_device()
lib = _cabi.lib()
Ta = torch.rand(plan.shape, dtype=torch.float64, device="cuda")
To = torch.empty_like(Ta)
W = torch.rand(plan.shape, dtype=torch.float64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
lib.hs2_debug_phase_strided.argtypes = [ctypes.c_void_p, ctypes.c_int]
names = ["wait tile", "smem->regs", "sync+issue next", "forward", "sync+interface", "sync+backward", "store (+Tin)", "-", "loop top"]
for which in ("y", "z"):
    def run():
        if which == "y":
            _cabi.check(lib.hs2_sweep_y(plan._handle, W.data_ptr(), st))
        else:
            _cabi.check(lib.hs2_sweep_z(plan._handle, Ta.data_ptr(), To.data_ptr(), W.data_ptr(), st))
    for it in range(2):
        run()
    torch.cuda.synchronize()
    lib.hs2_debug_phase_strided(None, 1)
    n = 3
    for it in range(n):
        run()
    torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 16)()
    lib.hs2_debug_phase_strided(out, 0)
    ntiles = 16384 * n
    tot = sum(out[:9])
    print(which, "sweep: cycles per tile (thread 0 of each block)")
    for i, nm in enumerate(names):
        if nm != "-":
            print("   %-16s %8.0f  %5.1f%%" % (nm, out[i] / ntiles, 100.0 * out[i] / max(tot, 1)))
    print("   total %.0f" % (tot / ntiles))

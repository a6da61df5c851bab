"""The fused slab z kernel WITHOUT a peer: one GPU, the other slabs' interface rows pre-filled (never empty), peers'
stores into a local dummy buffer or none - what does the kernel cost when it never waits?
    python scripts/zfused_bench.py [world] [rank] [steps]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import heatsim2_b200 as hs
from heatsim2_b200 import _cabi, crank_nicolson
from heatsim2_b200.plan import AdiPlan
import problems
import bench

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
shape = bench.grid_for(world, 512)
dev = torch.device("cuda", 0)
prob = problems.uniform_slab(hs, shape=shape, random_T0=False)
a = prob["setup_args"]
nz, ny, nx = shape
h = nz // world
k0 = rank * h
class_id, coefs, volume_array, vol = crank_nicolson.compile_problem(*a, device=dev)
plan = AdiPlan((h, ny, nx), None, coefs, a[9], volume_array, volumetric_elements=vol[k0:k0 + h], materials=a[10],
               slab=(k0, class_id))
del class_id
plan.ensure_device(dev)
lib = _cabi.lib()
M, Pg = plan.chunk[2]
p_loc = h // M
T = torch.rand((h, ny, nx), dtype=torch.float64, device=dev)
Tout = torch.empty_like(T)
work = torch.rand((h, ny, nx), dtype=torch.float64, device=dev)
n_lines = ny * nx
Yall = torch.rand((2 * Pg, n_lines), dtype=torch.float64, device=dev) * 1e-3
Ysave = Yall.clone()
dummy = torch.zeros((2 * p_loc, n_lines), dtype=torch.float64, device=dev)
status = torch.zeros(4, dtype=torch.int32, device=dev)
own = Yall[rank * 2 * p_loc:(rank + 1) * 2 * p_loc]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
H = plan._handle


def two(ev):
    ev[0].record()
    _cabi.check(lib.hs2_sweep_z_forward(H, work.data_ptr(), own.data_ptr(), 0, n_lines, st))
    _cabi.check(lib.hs2_sweep_z_backward(H, T.data_ptr(), Tout.data_ptr(), work.data_ptr(), Yall.data_ptr(), 0, n_lines, st))
    ev[1].record()


def fused(ev, n_peers):
    Yall.copy_(Ysave)          # the kernel empties the peers' slots it read
    arr = (ctypes.c_uint64 * 6)(*([dummy.data_ptr()] * 6))
    ev[0].record()
    _cabi.check(lib.hs2_sweep_z_fused(H, T.data_ptr(), Tout.data_ptr(), work.data_ptr(), Yall.data_ptr(), n_peers, arr, 5.0,
                                      status.data_ptr(), st))
    ev[1].record()


for name, fn in (("forward + backward launches", lambda e: two(e)), ("fused, no peer stores", lambda e: fused(e, 0)),
                 ("fused, 1 peer (stores into local memory)", lambda e: fused(e, 1)),
                 ("fused, 2 peers (local)", lambda e: fused(e, 2))):
    for _ in range(2):
        fn([torch.cuda.Event(enable_timing=True) for _ in range(2)])
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    for e in evs:
        fn(e)
    torch.cuda.synchronize()
    print("slab %dx%dx%d chunk %d x %d local: %-42s %.4f ms  (status %d)"
          % (h, ny, nx, M, p_loc, name, sum(e[0].elapsed_time(e[1]) for e in evs) / steps, int(status[0])))

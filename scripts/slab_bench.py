"""Time the kernels one rank of an N-GPU z-slab run launches per step, on ONE
GPU (no communication): x- and y-sweep of the slab, z_forward, z_backward.
    python scripts/slab_bench.py [world] [rank] [steps]
Grid = bench.py's weak-scaling grid for `world` GPUs."""
import ctypes
import os
import sys
import time

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
work = torch.empty_like(T)
halo = torch.rand((2, ny, nx), dtype=torch.float64, device=dev)
n_lines = ny * nx
Yall = torch.zeros((2 * Pg, n_lines), dtype=torch.float64, device=dev)
own = Yall[rank * 2 * p_loc:(rank + 1) * 2 * p_loc]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
H = plan._handle


def step(ev=None):
    if ev: ev[0].record()
    _cabi.check(lib.hs2_sweep_x(H, T.data_ptr(), work.data_ptr(), None, halo[0].data_ptr(), halo[1].data_ptr(), st))
    if ev: ev[1].record()
    _cabi.check(lib.hs2_sweep_y(H, work.data_ptr(), st))
    if ev: ev[2].record()
    _cabi.check(lib.hs2_sweep_z_forward(H, work.data_ptr(), own.data_ptr(), 0, n_lines, st))
    if ev: ev[3].record()
    _cabi.check(lib.hs2_sweep_z_backward(H, T.data_ptr(), Tout.data_ptr(), work.data_ptr(), Yall.data_ptr(), 0, n_lines, st))
    if ev: ev[4].record()


for _ in range(3):
    step()
torch.cuda.synchronize()
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(steps)]
for e in evs:
    step(e)
torch.cuda.synchronize()
names = ["x", "y", "z_fwd", "z_bwd"]
ms = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) / steps for i in range(4)]
print("slab %dx%dx%d (rank %d of %d, chunk %d, %d local chunks):" % (h, ny, nx, rank, world, M, p_loc),
      " ".join("%s %.4f" % (n, m) for n, m in zip(names, ms)), "total %.4f ms" % sum(ms),
      "exchange bytes/rank out: halo %d, interface %d per peer" % (2 * ny * nx * 8, 2 * p_loc * n_lines * 8))

"""Run an N-slab (multi-GPU style) ADI step sequence on ONE GPU: every slab's
kernels are launched one after the other through the C ABI and the two
exchanges (halo planes, z-interface values) are plain device copies.  TEST
INFRASTRUCTURE: gives the slab kernels (hs2_sweep_x with halos,
hs2_sweep_z_forward / hs2_sweep_z_backward) parity coverage on a single-GPU box."""
import ctypes

import numpy as np
import torch


def run(hs, prob, world, nsteps):
    from heatsim2_b200 import _cabi, crank_nicolson
    from heatsim2_b200.plan import AdiPlan
    dev = torch.device("cuda", torch.cuda.current_device())
    a = prob["setup_args"]
    nz, ny, nx = int(a[6]), int(a[7]), int(a[8])
    dt, materials, volumetric = a[9], a[10], a[12]
    assert nz % world == 0
    h = nz // world
    class_id, coefs, volume_array, vol = crank_nicolson.compile_problem(*a, device=dev)
    plans = []
    for r in range(world):
        slab_volume = volume_array[r * h:(r + 1) * h] if np.ndim(volume_array) > 0 else volume_array
        pl = AdiPlan((h, ny, nx), None, coefs, dt, slab_volume, volumetric_elements=vol[r * h:(r + 1) * h],
                     materials=materials, slab=(r * h, class_id))
        pl.ensure_device(dev)
        plans.append(pl)
    lib = _cabi.lib()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    M, Pg = plans[0].chunk[2]
    p_loc = h // M
    n_lines = ny * nx
    T = [torch.from_numpy(np.array(prob["T0"][r * h:(r + 1) * h])).to(dev) for r in range(world)]
    Tn = [torch.empty_like(t) for t in T]
    work = [torch.empty_like(t) for t in T]
    ve = prob["volumetric_elements"]
    for it in range(nsteps):
        t = prob["t0"] + it * prob["dt"]
        Yall = torch.zeros((2 * Pg, n_lines), dtype=torch.float64, device=dev)
        for r, pl in enumerate(plans):
            src = keep = None
            if volumetric is not None and len(volumetric):
                src, keep = pl._source(t, dt, ve[r * h:(r + 1) * h], volumetric)
            lo = T[r - 1][-1].clone() if r > 0 else None
            hi = T[r + 1][0].clone() if r < world - 1 else None
            # the two-launch form the peer-memory transport uses: interior planes, then the boundary planes
            for part in (1, 2):
                _cabi.check(lib.hs2_sweep_x_part(pl._handle, T[r].data_ptr(), work[r].data_ptr(),
                                                 ctypes.byref(src) if src is not None else None,
                                                 lo.data_ptr() if lo is not None else None,
                                                 hi.data_ptr() if hi is not None else None, part, st))
            _cabi.check(lib.hs2_sweep_y(pl._handle, work[r].data_ptr(), st))
            own = Yall[r * 2 * p_loc:(r + 1) * 2 * p_loc]
            _cabi.check(lib.hs2_sweep_z_forward(pl._handle, work[r].data_ptr(), own.data_ptr(), 0, n_lines, st))
            torch.cuda.synchronize()
        for r, pl in enumerate(plans):
            _cabi.check(lib.hs2_sweep_z_backward(pl._handle, T[r].data_ptr(), Tn[r].data_ptr(), work[r].data_ptr(),
                                                 Yall.data_ptr(), 0, n_lines, st))
        torch.cuda.synchronize()
        T, Tn = Tn, T
    return torch.cat(T, dim=0).cpu().numpy()

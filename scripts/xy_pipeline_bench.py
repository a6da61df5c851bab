"""Round-2 experiment (NOT YET RUN ON A GPU - written after the round-1 GPU budget was spent):
hand the x-sweep's result to the y-sweep through L2 instead of HBM.

Both sweeps are plane-local and the y-sweep overwrites the increment in place, so if the y-sweep of a
group of S planes runs right behind their x-sweep the increment (2 MB per plane at 512^2) is still in
the 126 MB L2: it is then neither written to nor read from HBM (56 -> 40 B per cell-update).  This script
does that with the EXISTING kernels and entry points: one slab plan per group of S planes (the multi-GPU
machinery: hs2_sweep_x takes the neighbouring planes as halo pointers, hs2_sweep_y is slab-local), x-sweeps
on one stream, y-sweeps on a second one behind per-slab events (so the tail of one kernel overlaps the
head of the next), then the ordinary z-sweep.  It checks that the field equals the plain step and prints
the timings of both as JSON lines.

    python scripts/xy_pipeline_bench.py [grid=512] [reps=20] [S ...=8 16 32]
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import heatsim2_b200 as hs
from heatsim2_b200 import _cabi, crank_nicolson
from heatsim2_b200.plan import AdiPlan
import problems


def main():
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    sizes = [int(v) for v in sys.argv[3:]] or [8, 16, 32]
    dev = torch.device("cuda", 0)
    lib = _cabi.lib()
    prob = problems.uniform_slab(hs, shape=(grid, grid, grid), random_T0=False)
    a = prob["setup_args"]
    nz, ny, nx = int(a[6]), int(a[7]), int(a[8])
    class_id, coefs, volume_array, vol = crank_nicolson.compile_problem(*a, device=dev)
    full = AdiPlan((nz, ny, nx), class_id, coefs, a[9], volume_array, volumetric_elements=vol, materials=a[10])
    full.ensure_device(dev)
    g = torch.Generator(device=dev).manual_seed(1234)
    T = torch.rand((nz, ny, nx), dtype=torch.float64, device=dev, generator=g)
    work = torch.empty_like(T)
    ref = torch.empty_like(T)
    out = torch.empty_like(T)
    main_s = torch.cuda.current_stream(dev)
    sA, sB = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def raw(s):
        return ctypes.c_void_p(s.cuda_stream)

    def plain_step():
        _cabi.check(lib.hs2_step(full._handle, T.data_ptr(), ref.data_ptr(), work.data_ptr(), None, None, None, raw(main_s)))

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_s)
        for _ in range(reps):
            fn()
        e1.record(main_s)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    base_ms = timed(plain_step)
    print(json.dumps({"variant": "plain hs2_step", "ms_per_step": base_ms, "G_cell_updates_per_s": nz * ny * nx / base_ms / 1e6}),
          flush=True)
    cid_full = full.class_id            # device tensor [nz, ny, nx]
    plane_bytes = ny * nx * 8
    for S in sizes:
        if nz % S:
            continue
        # one plan per DISTINCT group of S planes (a uniform slab has three: first, interior, last)
        plans, index = [], []
        for k0 in range(0, nz, S):
            found = None
            for (q0, pl) in plans:
                same_pos = (q0 == 0) == (k0 == 0) and (q0 + S == nz) == (k0 + S == nz)
                if same_pos and torch.equal(cid_full[q0:q0 + S], cid_full[k0:k0 + S]) and \
                        (q0 == 0 or torch.equal(cid_full[q0 - 1], cid_full[k0 - 1])) and \
                        (q0 + S == nz or torch.equal(cid_full[q0 + S], cid_full[k0 + S])):
                    found = pl
                    break
            if found is None:
                sv = volume_array[k0:k0 + S] if np.ndim(volume_array) > 0 else volume_array
                found = AdiPlan((S, ny, nx), None, coefs, a[9], sv, volumetric_elements=vol[k0:k0 + S], materials=a[10],
                                slab=(k0, class_id))
                found.ensure_device(dev)
                plans.append((k0, found))
            index.append(found)
        events = [torch.cuda.Event() for _ in index]

        def piped_step():
            sA.wait_stream(main_s)
            sB.wait_stream(main_s)
            for n, pl in enumerate(index):
                k0 = n * S
                t_ptr = T.data_ptr() + k0 * plane_bytes
                w_ptr = work.data_ptr() + k0 * plane_bytes
                lo = T.data_ptr() + (k0 - 1) * plane_bytes if k0 > 0 else None
                hi = T.data_ptr() + (k0 + S) * plane_bytes if k0 + S < nz else None
                _cabi.check(lib.hs2_sweep_x(pl._handle, t_ptr, w_ptr, None, lo, hi, raw(sA)))
                events[n].record(sA)
                sB.wait_event(events[n])
                _cabi.check(lib.hs2_sweep_y(pl._handle, w_ptr, raw(sB)))
            main_s.wait_stream(sB)
            _cabi.check(lib.hs2_sweep_z(full._handle, T.data_ptr(), out.data_ptr(), work.data_ptr(), raw(main_s)))

        plain_step()
        piped_step()
        torch.cuda.synchronize()
        err = float((out - ref).abs().max() / ref.abs().max())
        ms = timed(piped_step)
        print(json.dumps({"variant": "x/y pipelined through L2", "planes_per_group": S, "distinct_plans": len(plans),
                          "ms_per_step": ms, "G_cell_updates_per_s": nz * ny * nx / ms / 1e6, "relerr_vs_plain": err,
                          "bitwise_equal": bool(torch.equal(out, ref))}), flush=True)
        del plans, index


if __name__ == "__main__":
    main()

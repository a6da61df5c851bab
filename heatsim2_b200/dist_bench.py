"""bench.py's N>1 arm: weak-scaling z-slab run, one rank per GPU (launched by
torch.distributed.run).  Device-side timing, max over ranks, rank 0 prints."""
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

METRIC = "ADI cell-updates/s (float64)"
UNIT = "cell-updates/s"


def field_by_global_index(k0, k1, ny, nx, device, seed=1234):
    """T0[k,j,i] in [0,1) as a function of the GLOBAL flat cell index only (SURVEY 8d C5:
    rank-independent generation), so that an N-GPU run and a 1-GPU run of the same grid start
    from the same field.  splitmix64-style integer mix evaluated with wrapping int64 arithmetic
    on the device, one plane block at a time."""
    out = torch.empty((k1 - k0, ny, nx), dtype=torch.float64, device=device)
    plane = ny * nx
    m53 = (1 << 53) - 1

    def lsr(x, s):       # logical shift right of an int64 bit pattern
        return (x >> s) & ((1 << (64 - s)) - 1)

    def c(v):            # 64-bit constant as a signed python int
        return v - (1 << 64) if v >= (1 << 63) else v
    step = max(1, (1 << 24) // plane)
    for a in range(k0, k1, step):
        b = min(k1, a + step)
        x = torch.arange(a * plane, b * plane, dtype=torch.int64, device=device) + c((0x9E3779B97F4A7C15 * (seed + 1)) & ((1 << 64) - 1))
        x = (x ^ lsr(x, 30)) * c(0xBF58476D1CE4E5B9)
        x = (x ^ lsr(x, 27)) * c(0x94D049BB133111EB)
        x = x ^ lsr(x, 31)
        out[a - k0:b - k0] = (lsr(x, 11) & m53).to(torch.float64).mul_(1.0 / (1 << 53)).view(b - a, ny, nx)
    return out


def _global_sum(T):
    """sum over all ranks' slabs, compensated: (sum of per-plane sums) in float64"""
    s = T.sum(dim=(1, 2)).sum().reshape(1)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(s)


def run(args, shape, workload_name, clock_sampler=None):
    import heatsim2_b200 as hs
    from heatsim2_b200 import _cabi, dist as hdist
    import problems
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    _cabi.lib()
    if getattr(args, "shape", ""):
        shape = tuple(int(v) for v in args.shape.split(","))      # global grid override (experiments)
    t_p = time.perf_counter()
    prob = problems.uniform_slab(hs, shape=shape, random_T0=False)      # the caller's arrays (reference API): numpy, global grid
    problem_s = time.perf_counter() - t_p
    t0 = time.perf_counter()
    P, S = hdist.setup(*prob["setup_args"])
    dplan = P.plan
    dplan.plan.ensure_device(dev)
    setup_s = time.perf_counter() - t0
    k0, k1 = P.slab
    n_local = (k1 - k0) * shape[1] * shape[2]
    n_global = shape[0] * shape[1] * shape[2]
    Ta = field_by_global_index(k0, k1, shape[1], shape[2], dev)
    Tb = torch.empty_like(Ta)
    ve = prob["volumetric_elements"][k0:k1]
    vol = prob["volumetric"]
    dt = prob["dt"]
    hs.run_adi_steps(P, S, 0.0, dt, Ta, ve, vol, out=Tb)        # flash step, outside the timed region
    Ta, Tb = Tb, Ta
    it = 1
    for _ in range(max(args.warmup, 3)):
        hs.run_adi_steps(P, S, it * dt, dt, Ta, ve, vol, out=Tb)
        Ta, Tb = Tb, Ta
        it += 1
    sum_before = _global_sum(Ta)
    sampler = None
    if clock_sampler is not None and rank == 0:      # nvidia-smi clocks / throttle reasons of rank 0's GPU during the timed region
        sampler = clock_sampler(local)
        sampler.start()
        time.sleep(0.3)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        hs.run_adi_steps(P, S, it * dt, dt, Ta, ve, vol, out=Tb)
        Ta, Tb = Tb, Ta
        it += 1
    e1.record()
    torch.cuda.synchronize()
    t_end = time.time()
    dist.barrier()
    clocks = sampler.stop(t_begin, t_end) if sampler is not None else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms) / args.steps
    dplan.check()
    # uniform material in an insulated box: a source-free step conserves sum(T) exactly
    sum_after = _global_sum(Ta)
    check = {"sum_T_before": sum_before, "sum_T_after": sum_after,
             "rel_drift": abs(sum_after - sum_before) / abs(sum_before), "steps": args.steps,
             "field": "T0 = mix64(global cell index), rank-independent"}
    check["conserved"] = check["rel_drift"] <= 1e-12
    # N <= 2 (or HS2_BENCH_CHECK_1GPU=1): the same grid on ONE GPU (this rank's) from the same field,
    # same number of steps; max relative difference on this rank's slab, max over ranks
    want_1gpu = os.environ.get("HS2_BENCH_CHECK_1GPU", "1" if world <= 2 else "0") == "1"
    if want_1gpu:
        n_cmp = 3
        P1, S1 = hs.setup(*prob["setup_args"])
        A = field_by_global_index(0, shape[0], shape[1], shape[2], dev)
        B = torch.empty_like(A)
        a, b = A[k0:k1].clone(), torch.empty_like(Ta)
        for s in range(n_cmp):
            hs.run_adi_steps(P1, S1, s * dt, dt, A, prob["volumetric_elements"], vol, out=B)
            A, B = B, A
            hs.run_adi_steps(P, S, s * dt, dt, a, ve, vol, out=b)
            a, b = b, a
        diff = ((a - A[k0:k1]).abs().max() / A.abs().max()).reshape(1)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        check["vs_1gpu_max_rel_diff"] = float(diff)
        check["vs_1gpu_steps"] = n_cmp
        check["vs_1gpu_ok"] = float(diff) <= 1e-13
        del P1, S1, A, B, a, b
        torch.cuda.empty_cache()
    # per-phase device times of a few extra steps (peer-memory transport only; not part of the timed region)
    phase_ms = None
    if dplan._px is not None:
        dplan.profile = []
        for _ in range(5):
            hs.run_adi_steps(P, S, it * dt, dt, Ta, ve, vol, out=Tb)
            Ta, Tb = Tb, Ta
            it += 1
        phase_ms = dplan.profile_ms()
        dplan.profile = None
    all_phases = [None] * world
    dist.all_gather_object(all_phases, phase_ms)
    # e2e: every step each rank uploads its slab from pinned host memory and reads the result back
    H_in = torch.empty(Ta.shape, dtype=torch.float64).pin_memory()
    H_out = torch.empty(Ta.shape, dtype=torch.float64).pin_memory()
    H_in.copy_(Ta)
    e2e_steps = max(3, min(args.steps, 5))
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(e2e_steps):
        Ta.copy_(H_in, non_blocking=True)
        hs.run_adi_steps(P, S, it * dt, dt, Ta, ve, vol, out=Tb)
        H_out.copy_(Tb, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        H_in, H_out = H_out, H_in
        it += 1
    e1.record()
    torch.cuda.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2) / e2e_steps
    finite = torch.isfinite(Ta).all().to(torch.int32)
    dist.all_reduce(finite, op=dist.ReduceOp.MIN)
    if rank == 0:
        peak = 6650.0
        pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
        peak_src = "fallback (B200_PROFILING.md)"
        if os.path.exists(pk):
            peak, peak_src = json.load(open(pk))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy), per GPU"
        value = n_global / (ms_step * 1e-3)
        bytes_cell = 64          # distributed z-sweep reads the increment twice: 16 + 16 + 8 + 24
        comm = dplan.comm_bytes_per_step()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": getattr(args, "scaling", "weak"), "check": check,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(tuple(shape)), "grid": list(shape), "cells": n_global,
                           "cells_per_gpu": n_local, "decomposition": "z-slabs, %d planes per GPU" % (k1 - k0),
                           "l2": "inputs larger than L2", "setup_s": setup_s, "problem_arrays_s": problem_s,
                           "finite": bool(int(finite))},
                "roofline": {"bound": "hbm", "kernel": "whole step (3 sweeps + z interface exchange)",
                             "achieved": 56 * n_local / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": 56 * n_local / (ms_step * 1e-3) / 1e9 / peak, "traffic": None,
                             "peak_source": peak_src,
                             "note": "algorithmic 56 B/cell-update; the slab z-sweep actually moves %d B/cell" % bytes_cell},
                "comm": {"halo_bytes_per_rank_per_step": comm["halo_send"],
                         "interface_bytes_sent_per_rank_per_step": comm["interface_send"],
                         "interface_exchange": comm["interface_mode"], "rank0_phase_ms": phase_ms,
                         "phase_ms_by_rank": all_phases,
                         "backend": ("CUDA IPC peer memory over NVLink (NCCL only for set-up)" if dplan._px is not None else
                                     "NCCL %s over NVLink" % ".".join(str(v) for v in torch.cuda.nccl.version()))},
                "cpu_baseline": None,
                "e2e": {"value": n_global / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_global * 8,
                        "d2h_bytes_per_step": n_global * 8, "ms_per_step": e2e_ms, "steps": e2e_steps,
                        "api": "per rank: pinned host slab -> device, heatsim2_b200.run_adi_steps (dist plan), device -> pinned host"},
                "gpu_launches": args.steps * dplan.launches_per_step() * world, "clocks": clocks}
        print(json.dumps(line))
    dplan.check()
    dplan.close()
    dist.destroy_process_group()

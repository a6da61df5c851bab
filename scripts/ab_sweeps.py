#!/usr/bin/env python
"""A/B of sweep-kernel variants inside ONE process on one GPU (boxes differ by a few
percent, so only numbers from the same run are compared).

    python scripts/ab_sweeps.py [--shape nz,ny,nx] [--reps 20] NAME=ENV1:VAL,ENV2:VAL ...

Each variant is a name and a set of environment switches read by plan.py / the
library (HS2_UTAB_AXES, HS2_SLOT_PERM, HS2_X_KERNEL, ...); prints the mean x / y / z
sweep time (CUDA events around the three C-ABI calls) of source-free steps and one
JSON line per variant."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="512,512,512")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--problem", default="uniform_slab")
    ap.add_argument("variants", nargs="+")
    args = ap.parse_args()
    import numpy as np
    import torch
    import heatsim2_b200 as hs
    import problems
    shape = tuple(int(v) for v in args.shape.split(","))
    if args.problem == "uniform_slab":
        prob = problems.uniform_slab(hs, shape=shape, random_T0=False)
    else:
        prob = problems.ALL[args.problem](hs, nz=shape[0], ny=shape[1], nx=shape[2])
    g = torch.Generator(device="cuda").manual_seed(1)
    T0 = torch.rand(shape, dtype=torch.float64, device="cuda", generator=g)
    ref = None
    for spec in args.variants:
        name, _, envs = spec.partition("=")
        changed = {}
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition(":")
            changed[k] = os.environ.get(k)
            os.environ[k] = v
        try:
            P, S = hs.setup(*prob["setup_args"])
            plan = P.plan
            plan.ensure_device(torch.device("cuda", 0))
            Ta, Tb = T0.clone(), torch.empty_like(T0)
            for _ in range(3):
                hs.run_adi_steps(P, S, prob["dt"], prob["dt"], Ta, prob["volumetric_elements"], prob["volumetric"], out=Tb)
                Ta, Tb = Tb, Ta
            if ref is None:
                ref = Ta.clone()
            same = bool(torch.equal(ref, Ta))
            err = float((ref - Ta).abs().max() / ref.abs().max())
            evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.reps)]
            for k in range(args.reps):
                plan.timed_sweeps(Ta, Tb, evs[k])
                Ta, Tb = Tb, Ta
            torch.cuda.synchronize()
            ms = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) / args.reps for i in range(3)]
            out = {"variant": name, "env": {k: os.environ[k] for k in changed}, "shape": list(shape), "x_ms": ms[0],
                   "y_ms": ms[1], "z_ms": ms[2], "step_ms": sum(ms), "kernels": plan.last_kernels(),
                   "bitwise_equal_to_first": same, "rel_diff_to_first": err}
            print("%-22s x %.4f  y %.4f  z %.4f  step %.4f  %s %s" % (name, ms[0], ms[1], ms[2], sum(ms),
                                                                      plan.last_kernels(), "bit-identical" if same else "rel diff %.1e" % err))
            print(json.dumps(out))
            del P, S, plan
            torch.cuda.empty_cache()
        finally:
            for k, v in changed.items():
                if v is None:
                    del os.environ[k]
                else:
                    os.environ[k] = v


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/*.npz from the UNMODIFIED
reference (built into oracle/_ref by oracle/build_ref.py from /root/reference).

Run in the build container (the GPU box has no /root/reference; it only reads
the committed .npz files):

    python oracle/build_ref.py && python tests/golden/make_golden.py

Each file holds, for one tests/problems.py configuration, the field after the
listed step counts (``T_<n>``), the whole-run history at the probe cells
(``probe_hist``, when the problem names probes) and ``sum_<n>``; inputs are NOT
stored - they are rebuilt from tests/problems.py with the recorded kwargs.

BIG cases (SURVEY.md 8(d) sizes: C2 at 64^3 and 128^3 for 1000 steps, C3 at
128^3, C4 at 64x128x128) store a compact digest instead of the field
(tests/util.py ``digest``): the field sub-sampled with stride ``sub`` in every
direction, its sum, four dot products with seeded random weight fields (an error
anywhere in the field moves them) and, for C4, the per-step history of
surface_temperature.insulating_z_min_surface_temperature at a few columns.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import problems      # noqa: E402
import ref_loader    # noqa: E402

# name -> (problem, kwargs, steps whose full field is stored)
CASES = {
    "c1_steelonfoam": ("steelonfoam", dict(nsteps=999), (1, 10, 100, 999)),
    "c2_uniform_32": ("uniform_slab", dict(n=32, nsteps=200), (1, 10, 200)),
    "c2_uniform_ragged": ("uniform_slab", dict(shape=(19, 33, 45), nsteps=20), (1, 20)),
    "c3_steelonwater": ("steelonwater", dict(nz=80, ny=20, nx=24, nsteps=100), (1, 10, 100)),
    "c4_composite": ("composite", dict(nz=32, ny=32, nx=48, ply=8, nsteps=100), (1, 10, 100)),
    "sources_demo": ("sources_demo", dict(), (1, 2, 3, 4, 5, 6)),
    "line_1d": ("uniform_slab", dict(shape=(1, 1, 40), nsteps=5), (1, 5)),
    "plane_2d": ("uniform_slab", dict(shape=(1, 24, 20), nsteps=5), (1, 5)),
    "c6_curved_plate": ("curved_plate", dict(nz=24, ny=20, nx=28, nsteps=12), (1, 3, 12)),
    "c7_curved_map": ("curved_map", dict(nz=8, ny=10, nx=12, nsteps=6), (1, 3, 6)),     # one curvature pair per (j, i)
}


# name -> (problem, kwargs, steps digested, sub-sampling stride, surface-history columns or None)
BIG = {
    "big_c2_uniform_64": ("uniform_slab", dict(n=64, nsteps=1000), (1, 100, 1000), 4, None),
    "big_c2_uniform_128": ("uniform_slab", dict(n=128, nsteps=1000), (1, 100, 1000), 8, None),
    "big_c3_steelonwater_128": ("steelonwater", dict(nz=128, ny=128, nx=128, nsteps=200), (1, 20, 200), 8, None),
    "big_c4_composite_64x128x128": ("composite", dict(nz=64, ny=128, nx=128, ply=8, nsteps=300), (1, 30, 300), 8,
                                    ((5, 7), (40, 64), (64, 64), (100, 31))),
}


def big(ref, only):
    import util
    for case, (pname, kwargs, steps, sub, cols) in BIG.items():
        if only and case not in only:
            continue
        prob = problems.ALL[pname](ref, **kwargs)
        P, S = ref_loader.quiet_setup(ref, *prob["setup_args"])
        T = np.array(prob["T0"], dtype=np.float64)
        out = {"meta": json.dumps({"problem": pname, "kwargs": kwargs, "steps": list(steps), "sub": sub,
                                   "surface_cols": cols})}
        hist = []
        for it in range(prob["nsteps"]):
            T = ref.run_adi_steps(P, S, prob["t0"] + prob["dt"] * it, prob["dt"], T,
                                  prob["volumetric_elements"], prob["volumetric"])
            if cols:
                surf = ref.surface_temperature.insulating_z_min_surface_temperature(T, prob["dz"])
                hist.append([surf[c] for c in cols])
            if (it + 1) in steps:
                for k, v in util.digest(T, sub).items():
                    out["%s_%d" % (k, it + 1)] = v
        if cols:
            out["surface_hist"] = np.array(hist)
        np.savez_compressed(os.path.join(HERE, case + ".npz"), **out)
        print(case, prob["shape"], prob["nsteps"], "steps", flush=True)


def main():
    ref = ref_loader.load()
    if ref is None:
        raise SystemExit("reference not built: run python oracle/build_ref.py first")
    only = set(sys.argv[1:])          # optional: names of the cases to (re)generate
    if "--big" in only or any(o in BIG for o in only):
        import heatsim2.surface_temperature  # noqa: F401  (not imported by the reference's __init__)
        big(ref, only - {"--big"})
        return
    for case, (pname, kwargs, steps) in CASES.items():
        if only and case not in only:
            continue
        prob = problems.ALL[pname](ref, **kwargs)
        P, S = ref_loader.quiet_setup(ref, *prob["setup_args"])
        T = np.array(prob["T0"], dtype=np.float64)
        out = {"meta": json.dumps({"problem": pname, "kwargs": kwargs, "steps": list(steps)})}
        probes = prob.get("probes")
        hist = []
        for it in range(prob["nsteps"]):
            T = ref.run_adi_steps(P, S, prob["t0"] + prob["dt"] * it, prob["dt"], T,
                                  prob["volumetric_elements"], prob["volumetric"])
            if probes:
                hist.append([T[p] for p in probes])
            if (it + 1) in steps:
                # the big C1 fields are stored for two steps only
                if case != "c1_steelonfoam" or (it + 1) in (10, 100):
                    out["T_%d" % (it + 1)] = T.copy()
                out["sum_%d" % (it + 1)] = T.sum()
                if probes:
                    out["probe_%d" % (it + 1)] = np.array([T[p] for p in probes])
        if probes:
            out["probe_hist"] = np.array(hist)
        np.savez_compressed(os.path.join(HERE, case + ".npz"), **out)
        print(case, prob["shape"], prob["nsteps"], "steps")


if __name__ == "__main__":
    main()

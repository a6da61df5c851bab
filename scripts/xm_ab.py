"""A/B of the two x-sweep kernels on one GPU (same process, same field):
kernels_xf.cu (fold) vs kernels_xm.cu (march), x-sweep alone and whole step,
plus the march kernel's plane-range length.  Writes gpurun_out/xm_ab.json.

    python scripts/xm_ab.py [grid] [reps]
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch

import heatsim2_b200 as hs
from heatsim2_b200 import _cabi
import problems


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    lib = _cabi.lib()
    out = {"grid": grid, "reps": reps, "runs": []}
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1234)
    T = torch.rand((grid, grid, grid), dtype=torch.float64, device=dev, generator=g)
    T2 = torch.empty_like(T)
    W = {}
    # (kernel, extra environment read at launch time)
    variants = [("fold", {}),
                ("march", {"HS2_XM_R": "8", "HS2_XM_KR": "32"}),
                ("march", {"HS2_XM_R": "4", "HS2_XM_KR": "32"}),
                ("march", {"HS2_XM_R": "4", "HS2_XM_KR": "16"})]
    for chunk in ("32", "16"):
        os.environ["HS2_CHUNK_X"] = chunk
        prob = problems.uniform_slab(hs, shape=(grid, grid, grid), random_T0=False)
        P, S = hs.setup(*prob["setup_args"])
        plan = P.plan
        for name, env in variants:
            for k in ("HS2_XM_R", "HS2_XM_KR", "HS2_XM_TAB"):
                os.environ.pop(k, None)
            os.environ.update(env)
            plan.flags = 2 if name == "march" else 0
            plan.release()
            plan.ensure_device(dev)
            work = plan._buf("work")
            st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            try:
                x_ms = timed(lambda: _cabi.check(lib.hs2_sweep_x(plan._handle, T.data_ptr(), work.data_ptr(), None, None, None, st)), reps)
                Wx = work.clone()
                step_ms = timed(lambda: _cabi.check(lib.hs2_step(plan._handle, T.data_ptr(), T2.data_ptr(), work.data_ptr(),
                                                                  None, None, None, st)), reps)
                rec = {"chunk_x": int(chunk), "kernel": name, "x_kernel": plan.x_kernel, "env": env, "x_ms": x_ms,
                       "step_ms": step_ms, "G_cell_updates_per_s": grid ** 3 / step_ms / 1e6}
                if name == "fold":
                    W[chunk] = Wx
                elif chunk in W:
                    rec["relerr_vs_fold"] = float((Wx - W[chunk]).abs().max() / W[chunk].abs().max())
                    rec["bitwise_equal"] = bool(torch.equal(Wx, W[chunk]))
                del Wx
            except Exception as exc:                      # keep going: the other variants still get measured
                rec = {"chunk_x": int(chunk), "kernel": name, "env": env, "error": str(exc)[:300]}
            print(json.dumps(rec), flush=True)
            out["runs"].append(rec)
        W.pop(chunk, None)
        del plan, P, S
    ok = [r for r in out["runs"] if r.get("kernel") == "march" and "x_ms" in r and r.get("relerr_vs_fold", 1.0) <= 1e-14]
    folds = [r for r in out["runs"] if r.get("kernel") == "fold" and "x_ms" in r]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if ok:
        best = min(ok, key=lambda r: r["step_ms"])
        out["best_march"] = best
        out["best_fold"] = min(folds, key=lambda r: r["step_ms"]) if folds else None
        if out["best_fold"] is None or best["step_ms"] < out["best_fold"]["step_ms"]:
          with open(os.path.join(ROOT, "gpurun_out", "xm_best.env"), "w") as f:
            f.write("export HS2_X_KERNEL=march HS2_CHUNK_X=%d %s\n" % (best["chunk_x"], " ".join("%s=%s" % kv for kv in best["env"].items())))
        print("best march:", json.dumps(best))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "xm_ab.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

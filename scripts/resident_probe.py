"""Where the wall-clock time of run_adi_steps_n(numpy in, numpy out) goes (bench.py's e2e_resident)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import heatsim2_b200 as hs
import problems
g = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prob = problems.uniform_slab(hs, shape=(g, g, g), random_T0=False)
P, S = hs.setup(*prob["setup_args"])
plan = P.plan
plan.ensure_device()
A = np.random.default_rng(0).random((g, g, g))
def tic():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = tic(); cur = torch.empty(plan.shape, dtype=torch.float64, device=plan._dev); nxt = torch.empty_like(cur)
    t1 = tic(); plan.upload(A, cur)
    t2 = tic(); plan.run_steps_device(cur, nxt, 2, 50)
    t3 = tic(); out = plan.download(cur)
    t4 = tic()
    print("rep %d: alloc %.1f ms, upload (pageable) %.1f ms, 50 steps %.1f ms, download %.1f ms" % (rep, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3), flush=True)
    t5 = tic(); T2, rec = hs.run_adi_steps_n(P, S, 0.02, prob["dt"], out, prob["volumetric_elements"], prob["volumetric"], 50, probes=[(0, 1, 1)])
    t6 = tic()
    print("        run_adi_steps_n(pinned numpy in) 50 steps: %.1f ms" % ((t6-t5)*1e3), flush=True)
    del out, T2, cur, nxt

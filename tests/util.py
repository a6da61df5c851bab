"""Helpers shared by the tests."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def relerr(a, b):
    """max|a-b| / max|b| - the tolerance definition of BASELINE.json / SURVEY.md 8(d)."""
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a - b).max() / np.abs(b).max())


def digest(T, sub):
    """Compact digest of a field (big golden cases): sub-sampled values, sum and
    four dot products with seeded uniform(-1,1) weight fields; ``wabs`` = sum of
    |w*T| of the first one sets the scale of the dot-product tolerance."""
    T = np.asarray(T)
    out = {"sub": T[::sub, ::sub, ::sub].copy(), "sum": T.sum(), "absmax": np.abs(T).max()}
    dots = []
    for s in range(4):
        w = np.random.default_rng(1000 + s).uniform(-1.0, 1.0, T.shape)
        dots.append(float((w * T).sum()))
        if s == 0:
            out["wabs"] = float(np.abs(w * T).sum())
    out["dots"] = np.array(dots)
    return out


def check_digest(T, z, step, sub, tol):
    """Assert that field ``T`` matches the digest stored for ``step``."""
    d = digest(T, sub)
    want = z["sub_%d" % step]
    scale = float(z["absmax_%d" % step])
    assert float(np.abs(d["sub"] - want).max()) <= tol * scale, ("sub", step, float(np.abs(d["sub"] - want).max()) / scale)
    assert abs(d["sum"] - float(z["sum_%d" % step])) <= tol * T.size * scale, ("sum", step)
    # a dot product of n terms each within tol*scale of the reference
    lim = tol * scale * np.sqrt(T.size) * 4.0
    assert float(np.abs(d["dots"] - z["dots_%d" % step]).max()) <= lim, ("dots", step, float(np.abs(d["dots"] - z["dots_%d" % step]).max()) / lim)


def load_golden(case):
    z = np.load(os.path.join(GOLDEN, case + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith("big_"))


def big_golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith("big_"))


def run_b200(hs, prob, nsteps=None, record=None, host_api=False):
    """Run a tests/problems.py problem through heatsim2_b200.  Device-resident
    loop by default; host_api=True goes numpy -> run_adi_steps -> numpy."""
    import torch
    P, S = hs.setup(*prob["setup_args"])
    n = prob["nsteps"] if nsteps is None else nsteps
    rec = {}
    T = np.array(prob["T0"], dtype=np.float64)
    if not host_api:
        T = torch.from_numpy(T).cuda()
    for it in range(n):
        T = hs.run_adi_steps(P, S, prob["t0"] + prob["dt"] * it, prob["dt"], T,
                             prob["volumetric_elements"], prob["volumetric"])
        if record is not None and (it + 1) in record:
            rec[it + 1] = T.cpu().numpy().copy() if not host_api else T.copy()
    Tn = T if host_api else T.cpu().numpy()
    return (Tn, rec) if record is not None else Tn

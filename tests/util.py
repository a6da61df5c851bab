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


def load_golden(case):
    z = np.load(os.path.join(GOLDEN, case + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


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

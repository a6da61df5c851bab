"""Parity of the CUDA path (through the C ABI) with the oracle and with the
golden vectors produced by the unmodified reference.  Needs a B200."""
import numpy as np
import pytest

import problems
import util

pytestmark = pytest.mark.gpu

# tolerances of BASELINE.json: 1e-12 relative per step, 1e-10 after 1000 steps
TOL_STEP = 1e-12
TOL_RUN = 1e-10


@pytest.fixture(scope="module")
def hs():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import heatsim2_b200
    from heatsim2_b200 import _cabi
    _cabi.lib()          # fail loudly if the CUDA library is missing
    return heatsim2_b200


@pytest.mark.parametrize("case", util.golden_cases())
def test_golden_vectors(hs, case):
    z, meta = util.load_golden(case)
    prob = problems.ALL[meta["problem"]](hs, **meta["kwargs"])
    steps = meta["steps"]
    T, rec = util.run_b200(hs, prob, nsteps=max(steps), record=steps)
    for n in steps:
        tol = TOL_STEP if n == 1 else TOL_RUN
        if "T_%d" % n in z:
            assert util.relerr(rec[n], z["T_%d" % n]) <= tol, (case, n)
        assert abs(rec[n].sum() - float(z["sum_%d" % n])) <= tol * abs(float(z["sum_%d" % n])) * 10, (case, n)
        if "probe_%d" % n in z:
            got = np.array([rec[n][tuple(p)] for p in prob["probes"]])
            assert util.relerr(got, z["probe_%d" % n]) <= tol, (case, n)


@pytest.mark.parametrize("name,kwargs", [
    ("steelonfoam", dict(nsteps=5)),
    ("uniform_slab", dict(n=48, nsteps=5)),
    ("uniform_slab", dict(shape=(7, 130, 257), nsteps=3)),
    ("steelonwater", dict(nz=40, ny=40, nx=48, nsteps=5)),
    ("composite", dict(nz=32, ny=64, nx=64, nsteps=5)),
    ("sources_demo", dict()),
    ("curved_plate", dict()),
])
def test_single_steps_vs_oracle(hs, name, kwargs):
    """Every step restarted from the oracle's state: <= 1e-12."""
    import torch
    import adi_oracle
    prob = problems.ALL[name](hs, **kwargs)
    O = adi_oracle.setup(*prob["setup_args"])
    P, S = hs.setup(*prob["setup_args"])
    T = np.array(prob["T0"])
    for it in range(prob["nsteps"]):
        t = prob["t0"] + prob["dt"] * it
        want = O.step(t, prob["dt"], T)
        got = hs.run_adi_steps(P, S, t, prob["dt"], torch.from_numpy(T).cuda(),
                               prob["volumetric_elements"], prob["volumetric"]).cpu().numpy()
        assert util.relerr(got, want) <= TOL_STEP, (name, it)
        T = want


def test_host_api_matches_device_api(hs):
    prob = problems.steelonfoam(hs, nsteps=3)
    a = util.run_b200(hs, prob, host_api=True)
    b = util.run_b200(hs, prob, host_api=False)
    assert isinstance(a, np.ndarray)
    assert np.array_equal(a, b)


def test_energy_conservation_and_fixed_cells(hs):
    """All-insulated box: sum(rho c T dV) equals the flash energy; FIXED cells
    never change (crank_nicolson.pyx:473-486)."""
    prob = problems.steelonfoam(hs, nsteps=20)
    T = util.run_b200(hs, prob)
    rhoc = np.array([m[2] * m[3] if m[0] == 0 else 0.0 for m in prob["materials"]])[prob["material_elements"]]
    E = (rhoc * T).sum() * prob["dz"]            # J/m^2 per unit area, averaged over the face
    E /= (prob["shape"][1] * prob["shape"][2])
    assert abs(E - 10e3) <= 1e-9 * 10e3
    prob = problems.steelonwater(hs, nsteps=10)
    T = util.run_b200(hs, prob)
    fixed = prob["material_elements"] == 2
    assert np.array_equal(T[fixed], prob["T0"][fixed])

"""The drop-in boundary under the reference's own module names:
``dropin/heatsim2`` exposes heatsim2.alternatingdirection_c_pyx / crank_nicolson / tridiag on the B200 backend
and loads the UNCHANGED reference files (expression.py, boundary_*.py, surface_temperature.py, hs2_indexing.py)
from a reference tree when there is one; the reference's plug-in module objects go through ``setup``."""
import os
import subprocess
import sys

import numpy as np
import pytest

import problems
import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = ["/root/reference/heatsim2", os.path.join(ROOT, "oracle", "_ref", "heatsim2")]


def _ref_dir():
    for d in REF_DIRS:
        if os.path.isfile(os.path.join(d, "expression.py")):
            return d
    return None


def _run(code, ref=True, timeout=600):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT, os.path.join(ROOT, "tests"),
                                         os.path.join(ROOT, "oracle")])
    if ref and _ref_dir():
        env["HEATSIM2_REFERENCE"] = _ref_dir()
    else:
        env.pop("HEATSIM2_REFERENCE", None)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=timeout, cwd="/tmp")
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_module_names_of_the_reference_import():
    out = _run("""
import heatsim2, heatsim2.alternatingdirection_c_pyx, heatsim2.crank_nicolson, heatsim2.tridiag
import heatsim2.expression, heatsim2.boundary_conducting, heatsim2.surface_temperature, heatsim2.hs2_indexing
assert heatsim2.alternatingdirection is heatsim2.alternatingdirection_c_pyx
assert heatsim2.setup is heatsim2.crank_nicolson.setup and heatsim2.run_adi_steps is heatsim2.alternatingdirection_c_pyx.run_adi_steps
for n in ("run_adi_steps", "adi_setup", "adi_expressions", "add_equation_to_adi_matrices", "pyadi_step", "adi_params"):
    assert hasattr(heatsim2.alternatingdirection_c_pyx, n), n
for n in ("setup", "shift_expression", "subst_thermal_conductivity"):
    assert hasattr(heatsim2.crank_nicolson, n), n
assert hasattr(heatsim2.tridiag, "tridiaglu") and hasattr(heatsim2.tridiag, "tridiagsolve")
print(heatsim2.reference_dir, heatsim2.expression.__file__)
""", ref=False)
    assert "heatsim2_b200" in out          # without a reference tree the definition API is ours


@pytest.mark.skipif(_ref_dir() is None, reason="no reference tree here")
def test_unchanged_reference_files_are_used_and_their_plugins_pass_through_setup():
    """same equation classes whether the problem is defined with the reference's plug-in module objects (reference
    expression engine underneath) or with ours"""
    out = _run("""
import numpy as np, heatsim2, heatsim2_b200, problems
assert heatsim2.reference_dir is not None
assert "heatsim2_b200" not in heatsim2.expression.__file__ and "heatsim2_b200" not in heatsim2.boundary_conducting.__file__
assert heatsim2.boundary_conducting.group.__module__ == "heatsim2.expression"
assert hasattr(heatsim2.expression.linear_expression("T555"), "le_cmdlist")        # the reference's RPN engine
for name, kw in (("steelonfoam", dict(nz=16, ny=10, nx=12)), ("steelonwater", dict(nz=16, ny=10, nx=12)),
                 ("composite", dict(nz=16, ny=8, nx=8, ply=4))):
    a = problems.ALL[name](heatsim2, **kw)          # plug-ins = the reference's own module objects
    b = problems.ALL[name](heatsim2_b200, **kw)
    Pa, Sa = heatsim2.setup(*a["setup_args"])
    Pb, Sb = heatsim2_b200.setup(*b["setup_args"])
    ca = Pa.plan.class_coef[Pa.plan.class_id.numpy().astype(np.int64) & 0xFFFF]
    cb = Pb.plan.class_coef[Pb.plan.class_id.numpy().astype(np.int64) & 0xFFFF]
    err = np.abs(ca - cb).max() / np.abs(cb).max()
    assert err <= 1e-14, (name, err)
    print(name, "ok", err)
""")
    assert out.count("ok") == 3


def test_pyadi_step_add_equation_builds_the_same_plan():
    """reference method pyadi_step.add_equation (alternatingdirection_c_pyx.pyx:177-209), cell by cell"""
    import heatsim2_b200 as hs
    from heatsim2_b200 import alternatingdirection_c_pyx as adi
    prob = problems.steelonfoam(hs, nz=6, ny=5, nx=6)
    P, S = hs.setup(*prob["setup_args"])
    plan = P.plan
    cid = plan.class_id.numpy().astype(np.int64) & 0xFFFF
    # stage dictionaries of every class, rebuilt from the class coefficients in the reference's variable names
    def dicts(c):
        M, gxm, gxp, gym, gyp, gzm, gzp, D = plan.class_coef[c]
        g = {"x": (gxm, gxp), "y": (gym, gyp), "z": (gzm, gzp)}
        names = {"x": ("T554", "T556"), "y": ("T545", "T565"), "z": ("T455", "T655")}
        out = []
        for s, ax in enumerate("xyz"):
            d = {"volumetric_source": D, "T555m": M, "T555p%d" % s: -M}
            for a2 in "xyz":
                lo, hi = g[a2]
                n_lo, n_hi = names[a2]
                imp = "xyz".index(a2) <= s
                if imp:      # Crank-Nicolson average: half old (m), half new (p<stage of that axis>)
                    st = "xyz".index(a2)
                    for nm, gv in ((n_lo, lo), (n_hi, hi)):
                        d[nm + "m"] = d.get(nm + "m", 0.0) + 0.5 * gv
                        d[nm + "p%d" % st] = d.get(nm + "p%d" % st, 0.0) + 0.5 * gv
                    d["T555m"] -= 0.5 * (lo + hi)
                    d["T555p%d" % st] = d.get("T555p%d" % st, 0.0) - 0.5 * (lo + hi)
                else:
                    d[n_lo] = lo
                    d[n_hi] = hi
                    d["T555"] = d.get("T555", 0.0) - (lo + hi)
            out.append({k: v for k, v in d.items() if v != 0.0 or k in ("T555m",)})
        return out
    ADI_params, ADI_steps = adi.adi_setup(plan.shape, plan.volume_array)
    nz, ny, nx = plan.shape
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                ds = dicts(cid[k, j, i])
                for s in range(3):
                    ADI_steps[s].add_equation((k, j, i), ds[s])
    for st in ADI_steps:
        st.finalize()
    got = ADI_params.plan
    ca = got.class_coef[got.class_id.numpy().astype(np.int64) & 0xFFFF]
    cb = plan.class_coef[cid]
    assert np.abs(ca - cb).max() <= 1e-12 * np.abs(cb).max()


def test_hs2_indexing_round_trip():
    from heatsim2_b200 import hs2_indexing as ix
    shape = (6, 16, 32)
    scal = [0, 0, 1, 1, 2, 2]
    seen = set()
    for k in range(shape[0]):
        for j in range(shape[1]):
            for i in range(shape[2]):
                c = ix.compressed_index((k, j, i), scal)
                s = ix.compressed_single_index(c, scal, shape)
                assert ix.compressed_index_from_single(s, scal, shape) == c
                (k0, k1), (j0, j1), (i0, i1) = ix.uncompressed_index_range(c, scal)
                assert k0 <= k < k1 and j0 <= j < j1 and i0 <= i < i1
                seen.add(s)
    assert seen == set(range(2 * 16 * 32 + 2 * 8 * 16 + 2 * 4 * 8))


@pytest.mark.gpu
@pytest.mark.skipif(_ref_dir() is None, reason="no reference tree here")
def test_steelonfoam_demo_logic_through_the_dropin_names():
    """demos/steelonfoam.py's logic, unchanged but for the backend: ``import heatsim2`` resolves to the drop-in, the
    problem is built with the REFERENCE's plug-in objects, numpy in / numpy out; C1 golden (unmodified reference)"""
    out = _run("""
import numpy as np, heatsim2, problems, util
prob = problems.steelonfoam(heatsim2, nsteps=100)
(ADI_params, ADI_steps) = heatsim2.setup(*prob["setup_args"])
T = np.zeros((prob["nsteps"] + 1,) + prob["shape"], dtype='d')
for tcnt in range(prob["nsteps"]):
    t = prob["t0"] + prob["dt"] * tcnt
    T[tcnt + 1, ::] = heatsim2.run_adi_steps(ADI_params, ADI_steps, t, prob["dt"], T[tcnt, ::], prob["volumetric_elements"], prob["volumetric"])
z, meta = util.load_golden("c1_steelonfoam")
print("err100", util.relerr(T[100], z["T_100"]), "err10", util.relerr(T[10], z["T_10"]))
assert util.relerr(T[100], z["T_100"]) <= 1e-10 and util.relerr(T[10], z["T_10"]) <= 1e-10
print("demo ok")
""")
    assert "demo ok" in out

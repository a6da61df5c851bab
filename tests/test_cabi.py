"""The C-ABI library loads and exports every symbol include/hs2_b200.h declares
(no compute calls: this runs without a GPU)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "hs2_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hs2_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_header_symbols():
    from heatsim2_b200 import _cabi, build
    build.build()
    L = _cabi.lib()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_cabi.PROTOTYPES) == names
    assert L.hs2_abi_version() == _cabi.ABI_VERSION


def test_struct_layout_matches_header():
    import ctypes
    from heatsim2_b200 import _cabi
    # axis tables: 9 pointers + 6 int32; desc: int64 x3, int32 x2, ptr x2, axis[3], int32 x4
    assert ctypes.sizeof(_cabi.AxisTables) == 72 + 24
    assert ctypes.sizeof(_cabi.PlanDesc) == 24 + 8 + 16 + 3 * 96 + 16
    assert ctypes.sizeof(_cabi.Source) == 24
    # build desc: int64 x3, int32 x2, ptr x2, int32 x2, int32 x3 + int32, ptr, int64 x2
    assert ctypes.sizeof(_cabi.BuildDesc) == 24 + 8 + 16 + 8 + 16 + 8 + 16
    assert ctypes.sizeof(_cabi.AxisInfo) == 16 + 24
    L = _cabi.lib()
    for which, struct in enumerate((_cabi.AxisTables, _cabi.PlanDesc, _cabi.Source, _cabi.BuildDesc, _cabi.AxisInfo)):
        assert L.hs2_sizeof(which) == ctypes.sizeof(struct)


def test_argument_validation_without_gpu():
    import ctypes
    from heatsim2_b200 import _cabi
    L = _cabi.lib()
    out = ctypes.c_void_p()
    assert L.hs2_plan_create(None, ctypes.byref(out)) == -1
    assert b"NULL" in L.hs2_last_error()
    d = _cabi.PlanDesc()
    assert L.hs2_plan_create(ctypes.byref(d), ctypes.byref(out)) == -1
    assert b"empty grid" in L.hs2_last_error()
    assert L.hs2_tridiag_scratch_bytes(1 << 20) >= (1 << 20) // 2048 * 32
    assert L.hs2_step(None, None, None, None, None, None, None, None) == -1
    # many-steps / observation entry points (ABI 6): NULL plan or fields are argument errors, not crashes
    assert L.hs2_run_steps(None, None, None, None, 0, 10, 1, None, 0, None, None, 0.0, None, 1, None) == -1
    assert b"hs2_run_steps" in L.hs2_last_error()
    assert L.hs2_observe(None, None, None, 0, None, None, 0.0, None) == -1
    assert b"hs2_observe" in L.hs2_last_error()


def test_source_active_matches_evaluate_sources():
    """AdiPlan.source_active (time logic only; decides which steps run_adi_steps_n hands to hs2_run_steps) agrees with
    what evaluate_sources produces, step by step, on the problem that carries all four source kinds."""
    import numpy as np
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import heatsim2_b200 as hs
    import problems
    from heatsim2_b200.plan import AdiPlan
    prob = problems.sources_demo(hs)
    P, S = hs.setup(*prob["setup_args"])
    plan = P.plan
    active = []
    for n in range(12):
        t = prob["t0"] + n * prob["dt"]
        table, dense = plan.evaluate_sources(t, prob["dt"], prob["volumetric_elements"], prob["volumetric"])
        has = table is not None or (dense is not None and np.any(dense))
        assert AdiPlan.source_active(t, prob["volumetric"]) or not has       # never misses an active source
        active.append(AdiPlan.source_active(t, prob["volumetric"]))
    assert active[0] and not any(active[6:])         # the flash at t = 0; everything is over after 0.08 s
    assert not AdiPlan.source_active(0.5, None) and not AdiPlan.source_active(0.5, ())

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

"""hs2_plan_build (csrc/plan_build.cu) - the library's own table construction - against the numpy statement of the
same algebra in heatsim2_b200/plan.py.

CPU part: hs2_tables_chunk (the host algebra alone, no device) on random diagonally dominant lines and on the
lines of the test problems.  GPU part: every table of a natively built plan (hs2_plan_copy_table) against numpy,
the step results of the two table sources, the reference's out-of-domain error, and a plan driven from ctypes
alone (class ids + coefficient rows -> hs2_plan_build -> hs2_step) with no numpy table code in between."""
import ctypes

import numpy as np
import pytest

import problems
import util


def _native_chunk(lo, dg, hi, M, ghost=0):
    from heatsim2_b200 import _cabi
    L_ = _cabi.lib()
    nu, L = dg.shape
    P, pitch = -(-L // M), -(-L // 4) * 4
    tab = np.zeros((nu, 5, pitch))
    GE = np.zeros((nu, P, 2 * P))
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (lo, dg, hi)]
    band = L_.hs2_tables_chunk(*(a.ctypes.data_as(_cabi.c_double_p) for a in arrs), nu, L, M, ghost,
                               tab.ctypes.data_as(_cabi.c_double_p), GE.ctypes.data_as(_cabi.c_double_p))
    assert band >= 0, L_.hs2_last_error()
    return tab, GE, band


def _random_lines(rng, nu, L):
    lo = -rng.uniform(0.05, 3.0, (nu, L))
    hi = -rng.uniform(0.05, 3.0, (nu, L))
    lo[:, 0] = 0.0
    hi[:, -1] = 0.0
    # rows of I - (dt/2) M^-1 L: diagonal = 1 - lo - hi
    dg = 1.0 - lo - hi
    return lo, dg, hi


@pytest.mark.parametrize("L,M", [(48, 8), (130, 16), (257, 32), (512, 16), (512, 32), (1024, 32), (20, 8)])
def test_host_algebra_matches_numpy(L, M):
    from heatsim2_b200.plan import chunk_factors, interface_band
    rng = np.random.default_rng(L * 100 + M)
    lo, dg, hi = _random_lines(rng, 5, L)
    tab, GE, band = _native_chunk(lo, dg, hi, M)
    tab_np, GE_np = chunk_factors(lo, dg, hi, M)
    # the chunk-local recurrences are the same sequence of IEEE operations
    assert np.array_equal(tab, tab_np)
    # the inverse of the reduced system: Gauss-Jordan here, LAPACK there
    assert np.abs(GE - GE_np).max() <= 1e-13 * np.abs(GE_np).max()
    assert abs(band - interface_band(GE_np)) <= 1


def test_host_algebra_solves_the_line():
    """the partitioned solve with the native tables reproduces a dense solve (tests/emul.py statement of the kernels)"""
    import emul
    rng = np.random.default_rng(7)
    L, M = 200, 16
    lo, dg, hi = _random_lines(rng, 1, L)
    tab, GE, band = _native_chunk(lo, dg, hi, M)
    A = np.diag(dg[0]) + np.diag(lo[0, 1:], -1) + np.diag(hi[0, :-1], 1)
    d = rng.standard_normal((3, L))                            # three right hand sides on the same line
    want = np.linalg.solve(A, d.T).T
    P = -(-L // M)
    tabs = np.repeat(tab[:, :, :], 3, axis=0)                  # [lines][5][pitch]
    u, Y = emul.chunk_forward(d, tabs, M)
    E = Y @ GE[0].T                                            # [lines][P]
    alpha = np.hstack([np.zeros((3, 1)), E[:, :-1]])
    got = emul.chunk_backward(u, tabs, M, E, alpha)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


def test_ghost_tables_match_numpy():
    from heatsim2_b200.plan import ghost_uniform_tables
    L, M = 512, 16
    a, c = -0.37, -0.37
    b = 1.0 - a - c
    lo = np.full((1, L), a); dg = np.full((1, L), b); hi = np.full((1, L), c)
    tab, GE, band = _native_chunk(lo, dg, hi, M, ghost=1)
    lo2, dg2, hi2 = lo.copy(), dg.copy(), hi.copy()
    lo2[0, 0] = 0.0; hi2[0, -1] = 0.0; dg2[0, 0] = b + a; dg2[0, -1] = b + c
    code, xw, band_np = ghost_uniform_tables(lo2, dg2, hi2, M)
    assert code[0] == 1 and band == band_np
    P = L // M
    w = 2 * band + 1
    rel = xw[0, 128:].reshape(w, P, 2)
    for d in range(w):
        for p in range(P):
            q = p + d - band
            if 0 <= q < P:
                assert abs(rel[d, p, 0] - GE[0, p, 2 * q]) <= 1e-14
                assert abs(rel[d, p, 1] - GE[0, p, 2 * q + 1]) <= 1e-14


def test_bad_arguments_without_gpu():
    from heatsim2_b200 import _cabi
    L_ = _cabi.lib()
    out = ctypes.c_void_p()
    assert L_.hs2_plan_build(None, ctypes.byref(out)) == -1
    b = _cabi.BuildDesc()
    assert L_.hs2_plan_build(ctypes.byref(b), ctypes.byref(out)) == -1
    assert b"empty grid" in L_.hs2_last_error()
    assert L_.hs2_tables_chunk(None, None, None, 1, 8, 8, 0, None, None) == -1
    assert L_.hs2_plan_copy_table(None, 0, 0, None, 0) == -1


# ------------------------------------------------------------------------------------------------ GPU
CASES = [
    ("uniform_slab", dict(shape=(12, 20, 48))),
    ("uniform_slab", dict(shape=(4, 6, 1024))),
    ("steelonfoam", dict(nz=12, ny=20, nx=256, nsteps=2)),
    ("steelonwater", dict(nz=9, ny=14, nx=512)),
    ("composite", dict(nz=40, ny=24, nx=64)),
    ("steelonwater", dict(nz=8, ny=130, nx=16)),
    ("composite", dict(nz=257, ny=8, nx=16)),
]


@pytest.fixture(scope="module")
def hs():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import heatsim2_b200
    from heatsim2_b200 import _cabi
    _cabi.lib()
    return heatsim2_b200


@pytest.mark.gpu
@pytest.mark.parametrize("name,kwargs", CASES)
def test_native_tables_match_numpy(hs, name, kwargs):
    from heatsim2_b200 import _cabi
    from heatsim2_b200 import plan as planmod
    prob = problems.ALL[name](hs, **kwargs)
    P, S = hs.setup(*prob["setup_args"])
    plan = P.plan
    plan.ensure_device()
    assert plan._desc.__class__ is _cabi.BuildDesc            # the native builder made this plan
    chunk = plan.chunk
    for axis in range(3):
        L = plan.shape[2 - axis]
        info = plan.axis_info(axis)
        lo, dg, hi = plan.line_rows[axis]
        nu = dg.shape[0]
        assert info.n_unique == nu and info.line_length == L
        assert np.array_equal(plan.copy_table(axis, _cabi.TAB_LINE_ID, np.uint32), plan.line_id[axis].cpu().numpy().astype(np.uint32))
        for which, want in ((_cabi.TAB_ROWS_LO, lo), (_cabi.TAB_ROWS_DG, dg), (_cabi.TAB_ROWS_HI, hi)):
            assert np.array_equal(plan.copy_table(axis, which).reshape(nu, L), want)
        assert np.array_equal(plan.copy_table(axis, _cabi.TAB_LU).reshape(nu, L, 4), plan.line_lu[axis])
        M, Pn = chunk[axis]
        if not M:
            continue
        tab_np, GE_np = plan.chunk_tabs[axis]
        assert np.array_equal(plan.copy_table(axis, _cabi.TAB_CHUNK).reshape(tab_np.shape), tab_np)
        GE = plan.copy_table(axis, _cabi.TAB_GE).reshape(GE_np.shape)
        assert np.abs(GE - GE_np).max() <= 1e-13 * np.abs(GE_np).max()
        assert abs(info.band - planmod.interface_band(GE_np)) <= 1
        if axis == 0:
            il = planmod.interleave_chunks(tab_np, L, M, Pn)
            assert np.array_equal(plan.copy_table(0, _cabi.TAB_CHUNK_IL).reshape(il.shape), il)
        if "xyz"[axis] in plan.utab_axes:
            utab, ucode = planmod.uniform_chunks(tab_np, L, M, Pn, plan.line_id[axis])
            got_code = plan.copy_table(axis, _cabi.TAB_UCODE, np.uint8).reshape(nu, Pn)
            got_utab = plan.copy_table(axis, _cabi.TAB_UTAB).reshape(5, M)
            # the most common chunk may tie; both choices must be self-consistent
            if np.array_equal(got_utab, utab):
                assert np.array_equal(got_code, ucode)
            else:
                blocks = tab_np[:, :, :L // M * M].reshape(nu, 5, L // M, M)
                for u in range(nu):
                    for p in range(L // M):
                        assert bool(got_code[u, p]) == np.array_equal(blocks[u, :, p], got_utab)
        if axis == 0 and info.xw_band >= 0 and planmod.x_warp_applies(L):
            code, xw, band = planmod.ghost_uniform_tables(lo, dg, hi, 16)
            assert np.array_equal(plan.copy_table(0, _cabi.TAB_XW_CODE, np.uint8), code)
            if code.any():
                assert info.xw_band == band
                got = plan.copy_table(0, _cabi.TAB_XW).reshape(nu, -1)
                assert got.shape == xw.shape
                assert np.abs(got - xw).max() <= 1e-13 * np.abs(xw).max()


@pytest.mark.gpu
@pytest.mark.parametrize("name,kwargs", CASES[:5])
def test_native_and_numpy_tables_step_alike(hs, monkeypatch, name, kwargs):
    import adi_oracle
    prob = problems.ALL[name](hs, **kwargs)
    got_native = util.run_b200(hs, prob, nsteps=2)
    monkeypatch.setenv("HS2_TABLES", "numpy")
    got_numpy = util.run_b200(hs, prob, nsteps=2)
    assert util.relerr(got_native, got_numpy) <= 1e-13
    assert util.relerr(got_native, adi_oracle.run(prob, nsteps=2)) <= 2e-12


@pytest.mark.gpu
def test_ctypes_only_client(hs):
    """what a Cython / C maintainer would do: class ids + coefficient rows in, time steps out; no numpy table code"""
    import torch
    import adi_oracle
    from heatsim2_b200 import _cabi
    prob = problems.steelonwater(hs, nz=10, ny=18, nx=64)
    P, S = hs.setup(*prob["setup_args"])
    cid = P.plan.class_id.cuda().contiguous()
    coef = np.ascontiguousarray(P.plan.class_coef)
    L_ = _cabi.lib()
    b = _cabi.BuildDesc()
    b.nz, b.ny, b.nx = prob["shape"]
    b.n_classes = coef.shape[0]
    b.class_id_bytes = cid.element_size()
    b.d_class_id = cid.data_ptr()
    b.h_class_coef = coef.ctypes.data_as(_cabi.c_double_p)
    b.device = torch.cuda.current_device()
    b.utab_axes = 5
    h = ctypes.c_void_p()
    _cabi.check(L_.hs2_plan_build(ctypes.byref(b), ctypes.byref(h)))
    T = torch.from_numpy(np.array(prob["T0"])).cuda()
    out = torch.empty_like(T)
    work = torch.empty_like(T)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _cabi.check(L_.hs2_step(h, T.data_ptr(), out.data_ptr(), work.data_ptr(), None, None, None, st))
    torch.cuda.synchronize()
    # (a time at which no volumetric source fires: the client above passed none)
    want = adi_oracle.setup(*prob["setup_args"]).step(prob["t0"] + 0.5 * prob["dt"], prob["dt"], np.array(prob["T0"]))
    assert util.relerr(out.cpu().numpy(), want) <= 1e-12
    assert L_.hs2_plan_destroy(h) == 0


@pytest.mark.gpu
def test_open_face_is_an_error(hs):
    """a conductance pointing out of the grid: the reference exits in C add_equation (alternatingdirection_c.c:160-163)"""
    import torch
    from heatsim2_b200 import _cabi
    L_ = _cabi.lib()
    cid = torch.zeros((4, 5, 6), dtype=torch.uint8, device="cuda")
    coef = np.array([[1.0, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 1.0]])
    b = _cabi.BuildDesc()
    b.nz, b.ny, b.nx = 4, 5, 6
    b.n_classes, b.class_id_bytes = 1, 1
    b.d_class_id = cid.data_ptr()
    b.h_class_coef = coef.ctypes.data_as(_cabi.c_double_p)
    b.device = torch.cuda.current_device()
    h = ctypes.c_void_p()
    assert L_.hs2_plan_build(ctypes.byref(b), ctypes.byref(h)) == -1
    assert b"exceeds bounds of domain" in L_.hs2_last_error()

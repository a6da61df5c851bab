"""z-marching x-sweep kernel (csrc/kernels_xm.cu, HS2_FLAG_X_MARCH) against the
folded tile kernel (kernels_xf.cu) on the same field - the arithmetic is the
same, so full chunks must agree bit for bit - and against the oracle through
whole ADI steps."""
import ctypes

import numpy as np
import pytest

import problems
import util

pytestmark = pytest.mark.gpu

FLAG_MARCH = 2


@pytest.fixture(scope="module")
def hs():
    import heatsim2_b200
    from heatsim2_b200 import _cabi
    _cabi.lib()
    return heatsim2_b200


def _sweep_x(hs, prob, flags, T):
    import torch
    from heatsim2_b200 import _cabi
    P, S = hs.setup(*prob["setup_args"])
    plan = P.plan
    plan.flags = flags
    plan.ensure_device(T.device)
    W = torch.full_like(T, float("nan"))
    st = ctypes.c_void_p(torch.cuda.current_stream(T.device).cuda_stream)
    _cabi.check(_cabi.lib().hs2_sweep_x(plan._handle, T.data_ptr(), W.data_ptr(), None, None, None, st))
    torch.cuda.synchronize()
    return W.cpu().numpy(), plan


CASES = [
    ("steelonfoam", dict(), None, None),                              # C1 as shipped: 80x40x48, 2 line classes
    ("uniform_slab", dict(shape=(40, 72, 96)), None, None),
    ("uniform_slab", dict(shape=(33, 34, 50)), None, None),           # ragged chunks, ragged last tile
    ("uniform_slab", dict(shape=(9, 24, 512)), None, None),           # bench line length: 2 copy boxes, M=32
    ("uniform_slab", dict(shape=(70, 16, 300)), "32", None),          # second copy box partly out of range
    ("uniform_slab", dict(shape=(7, 40, 256)), "16", "3"),            # M=16, plane ranges 3+3+1
    ("uniform_slab", dict(shape=(12, 20, 64)), "8", "1"),             # M=8, one plane per item
    ("steelonwater", dict(nz=64, ny=48, nx=56), None, "5"),           # FIXED layer, thin layers, several line classes
    ("composite", dict(nz=32, ny=64, nx=64, ply=8), None, None),
]


@pytest.mark.parametrize("name,kwargs,chunk_x,kr", CASES)
@pytest.mark.parametrize("R", ["8", "4"])
def test_march_equals_fold(hs, name, kwargs, chunk_x, kr, R, monkeypatch):
    import torch
    monkeypatch.setenv("HS2_XM_R", R)                   # x-lines per tile
    if chunk_x:
        monkeypatch.setenv("HS2_CHUNK_X", chunk_x)
    if kr:
        monkeypatch.setenv("HS2_XM_KR", kr)
    prob = problems.ALL[name](hs, **kwargs)
    g = torch.Generator(device="cuda").manual_seed(7)
    shape = tuple(np.array(prob["T0"]).shape)
    T = torch.rand(shape, dtype=torch.float64, device="cuda", generator=g)
    a, plan_a = _sweep_x(hs, prob, 0, T)
    b, plan_b = _sweep_x(hs, prob, FLAG_MARCH, T)
    assert plan_a.x_kernel == "fold" and plan_b.x_kernel == "march"
    assert np.isfinite(b).all()
    assert util.relerr(b, a) <= 1e-14
    if shape[2] % plan_b.chunk[0][0] == 0:    # full chunks only: identical instruction sequence, identical bits
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name,kwargs,nsteps", [
    ("steelonfoam", dict(), 20),
    ("steelonwater", dict(nz=64, ny=48, nx=56), 6),
    ("uniform_slab", dict(shape=(33, 34, 50)), 6),
    ("sources_demo", dict(nz=12, ny=10, nx=14), None),     # steps with sources go through kernels_xf.cu, the others march
])
@pytest.mark.parametrize("R", ["8", "4"])
def test_march_steps_match_oracle(hs, name, kwargs, nsteps, R, monkeypatch):
    import adi_oracle
    monkeypatch.setenv("HS2_X_KERNEL", "march")
    monkeypatch.setenv("HS2_XM_R", R)
    prob = problems.ALL[name](hs, **kwargs)
    n = prob["nsteps"] if nsteps is None else nsteps
    got = util.run_b200(hs, prob, nsteps=n)
    assert util.relerr(got, adi_oracle.run(prob, nsteps=n)) <= 1e-12 * max(1, n // 10)


def test_more_than_256_classes_march(hs):
    """u16 class ids, coefficients read from global memory"""
    import torch
    rng = np.random.default_rng(5)
    nz, ny, nx = 12, 14, 16
    prob = problems.uniform_slab(hs, shape=(nz, ny, nx))
    args = list(prob["setup_args"])
    nmat = 6
    args[10] = tuple((hs.TEMPERATURE_COMPUTE, float(10 + 7 * m), 7.0e3 + 100 * m, 400.0 + 10 * m) for m in range(nmat))
    args[13] = rng.integers(0, nmat, size=(nz, ny, nx)).astype(np.uint8)
    prob = dict(prob, setup_args=tuple(args))
    T = torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda")
    a, plan_a = _sweep_x(hs, prob, 0, T)
    b, plan_b = _sweep_x(hs, prob, FLAG_MARCH, T)
    assert plan_b.n_classes > 256 and plan_b.x_kernel == "march"
    assert util.relerr(b, a) <= 1e-14


def test_small_grids_fall_back_to_fold(hs):
    prob = problems.uniform_slab(hs, shape=(6, 8, 16))      # fewer rows than a 10-row slice
    P, S = hs.setup(*prob["setup_args"])
    P.plan.flags = FLAG_MARCH
    assert P.plan.x_kernel == "fold"

"""CPU check of the z-marching x-sweep kernel's indexing and copy/wait protocol
(csrc/kernels_xm.cu) through its thread-level transcription tests/xm_sim.py:
the result must equal the plain stencil + chunked line solve of tests/emul.py,
every wait must find exactly one completed copy and every slice must hold the
plane the kernel believes it holds (asserted inside the simulation)."""
import numpy as np
import pytest

import emul
import problems
import xm_sim


def _plan(name, kwargs, chunk_x=None, monkeypatch=None):
    import heatsim2_b200 as hs
    if chunk_x is not None:
        monkeypatch.setenv("HS2_CHUNK_X", str(chunk_x))
    prob = problems.ALL[name](hs, **kwargs)
    P, S = hs.setup(*prob["setup_args"])
    return P.plan


@pytest.mark.parametrize("name,kwargs,KR,chunk_x", [
    ("steelonfoam", dict(nz=7, ny=20, nx=28), 32, None),          # one item per tile, ragged last tile (20 = 8+8+4)
    ("steelonfoam", dict(nz=9, ny=12, nx=28), 4, None),           # plane ranges 4+4+1, several line classes
    ("uniform_slab", dict(shape=(5, 10, 72)), 2, 32),             # M=32: two column pairs per thread, short last chunk
    ("uniform_slab", dict(shape=(3, 11, 264)), 1, 32),            # two copy boxes per row (nx > 256), KR = 1
    ("steelonwater", dict(nz=6, ny=20, nx=24), 3, 8),             # FIXED cells, thin layer
    ("composite", dict(nz=8, ny=16, nx=32, ply=4), 5, 16),
])
@pytest.mark.parametrize("R", [8, 4])
def test_march_kernel_transcription_equals_plain_sweep(name, kwargs, KR, chunk_x, R, monkeypatch):
    plan = _plan(name, kwargs, chunk_x, monkeypatch)
    rng = np.random.default_rng(3)
    T = rng.random(plan.shape)
    want = emul.solve_axis(plan, emul.rhs(plan, T, None, None, None), 0)
    got = xm_sim.sweep_x(plan, T, KR=KR, n_blocks=3, R=R)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


def test_geometry_limits():
    # slices of 10 rows: shorter grids, odd nx and lines needing more than 256 threads fall back
    assert xm_sim.geometry(8, 9, 32, 8, 4, 32) is None
    assert xm_sim.geometry(8, 16, 31, 8, 4, 32) is None
    assert xm_sim.geometry(8, 16, 1024, 32, 32, 32) is None
    g = xm_sim.geometry(512, 512, 512, 32, 16, 32)
    assert g["threads"] == 512 and g["RPT"] == 4 and g["NXB"] == 2 and g["n_items"] == 1024 and g["buf_alias"]
    # bytes of the four slices (the solve buffer lives in one of them) + tables at 512^3 stay inside 227 KB
    smem = (4 * g["slot_stride"] + 3 * 16 * 8 + 27 * 8 + 5 * 16 * 34 + 16 * 34) * 8 + 32
    assert smem <= 232448
    assert xm_sim.geometry(64, 64, 256, 16, 16, 32)["RPT"] == 2
    assert not xm_sim.geometry(12, 20, 64, 8, 8, 32)["buf_alias"]      # M=8: padded rows outgrow a slice

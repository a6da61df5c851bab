"""More GPU parity: tile kernels vs whole-line fallback, u16 class ids, odd
sizes, long runs, and size-independent properties at full BASELINE sizes."""
import numpy as np
import pytest

import problems
import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hs():
    import heatsim2_b200
    from heatsim2_b200 import _cabi
    _cabi.lib()
    return heatsim2_b200


def _run_plan(hs, prob, nsteps, flags=0):
    import torch
    P, S = hs.setup(*prob["setup_args"])
    P.plan.flags = flags
    T = torch.from_numpy(np.array(prob["T0"])).cuda()
    for it in range(nsteps):
        T = hs.run_adi_steps(P, S, prob["t0"] + it * prob["dt"], prob["dt"], T,
                             prob["volumetric_elements"], prob["volumetric"])
    return T.cpu().numpy(), P.plan


@pytest.mark.parametrize("name,kwargs", [
    ("steelonfoam", dict()),
    ("uniform_slab", dict(shape=(40, 72, 96))),
    ("uniform_slab", dict(shape=(33, 34, 50))),       # ragged chunks
    ("steelonwater", dict(nz=64, ny=48, nx=56)),
])
def test_tile_kernels_equal_fallback(hs, name, kwargs):
    prob = problems.ALL[name](hs, **kwargs)
    a, plan = _run_plan(hs, prob, 4, flags=0)
    b, _ = _run_plan(hs, prob, 4, flags=1)             # HS2_FLAG_FORCE_FALLBACK
    assert plan.launches_per_step == 3
    assert util.relerr(a, b) <= 1e-13


def test_odd_nx_uses_fallback_for_x_and_still_matches_oracle(hs):
    import adi_oracle
    prob = problems.uniform_slab(hs, shape=(9, 20, 31))
    got, plan = _run_plan(hs, prob, 3)
    assert plan.launches_per_step == 4                  # x: rhs + whole-line solve
    assert util.relerr(got, adi_oracle.run(prob, nsteps=3)) <= 1e-12


def test_more_than_256_classes_u16_ids(hs):
    """random 3-D material mosaic -> several hundred equation classes"""
    import adi_oracle
    rng = np.random.default_rng(5)
    nz, ny, nx = 12, 14, 16
    prob = problems.uniform_slab(hs, shape=(nz, ny, nx))
    args = list(prob["setup_args"])
    nmat = 6
    args[10] = tuple((hs.TEMPERATURE_COMPUTE, float(10 + 7 * m), 7.0e3 + 100 * m, 400.0 + 10 * m) for m in range(nmat))
    me = rng.integers(0, nmat, size=(nz, ny, nx)).astype(np.uint8)
    args[13] = me
    prob = dict(prob, setup_args=tuple(args), material_elements=me, materials=args[10])
    P, S = hs.setup(*prob["setup_args"])
    assert P.plan.n_classes > 256 and P.plan.class_id.element_size() == 2
    got, _ = _run_plan(hs, prob, 3)
    assert util.relerr(got, adi_oracle.run(prob, nsteps=3)) <= 1e-12


def test_c2_long_run_vs_oracle(hs):
    """C2 recipe at 48^3, 300 free-running steps (<= 1e-10)"""
    import adi_oracle
    prob = problems.uniform_slab(hs, n=48, nsteps=300)
    got, _ = _run_plan(hs, prob, 300)
    assert util.relerr(got, adi_oracle.run(prob)) <= 1e-10


def test_full_size_properties_256(hs):
    """256^3 (BASELINE configs[1] size): energy conservation in the insulated
    box, linearity of the step, finite output."""
    import torch
    prob = problems.uniform_slab(hs, n=256, random_T0=False)
    P, S = hs.setup(*prob["setup_args"])
    dt = prob["dt"]
    ve, vol = prob["volumetric_elements"], prob["volumetric"]
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.rand(P.plan.shape, dtype=torch.float64, device="cuda", generator=g)
    B = torch.rand(P.plan.shape, dtype=torch.float64, device="cuda", generator=g)

    def step(T, t=dt):
        return hs.run_adi_steps(P, S, t, dt, T, ve, vol)
    sA, sB = step(A), step(B)
    # uniform material: sum(T) is conserved by a source-free step
    assert abs(float(sA.sum() - A.sum())) <= 1e-12 * float(A.sum())
    lin = step(2.0 * A - 3.0 * B)
    assert float((lin - (2.0 * sA - 3.0 * sB)).abs().max()) <= 1e-12 * float(lin.abs().max())
    # the flash deposits exactly 10 kJ/m^2 on layer 0: rho c dz * sum over z = 1e4 per (y,x) column
    F = hs.run_adi_steps(P, S, 0.0, dt, torch.zeros_like(A), ve, vol)
    rho, c = prob["materials"][0][2], prob["materials"][0][3]
    E = float(F.sum()) * rho * c * prob["dz"] / (256 * 256)
    assert abs(E - 10e3) <= 1e-9 * 10e3
    assert bool(torch.isfinite(F).all())


def test_device_resident_loop_with_observation(hs):
    """run_adi_steps_n: probes and surface temperature recorded on the device
    equal the step-by-step host loop of demos/steelonfoam.py; C1 golden probes."""
    z, meta = util.load_golden("c1_steelonfoam")
    prob = problems.steelonfoam(hs, nsteps=100)
    P, S = hs.setup(*prob["setup_args"])
    Tn, rec = hs.run_adi_steps_n(P, S, prob["t0"], prob["dt"], prob["T0"], prob["volumetric_elements"],
                                 prob["volumetric"], 100, probes=prob["probes"], surface_dz=prob["dz"], every=1)
    assert isinstance(Tn, np.ndarray) and rec["probes"].shape == (100, 2) and rec["surface"].shape == (100, 40, 48)
    assert util.relerr(rec["probes"], z["probe_hist"][:100]) <= 1e-10
    assert util.relerr(Tn, z["T_100"]) <= 1e-10
    want_surface = hs.surface_temperature.insulating_z_min_surface_temperature(Tn, prob["dz"])
    assert util.relerr(rec["surface"][-1], want_surface) <= 1e-14


@pytest.mark.parametrize("every,nsteps,use_graph", [(1, 23, True), (3, 40, True), (4, 40, True), (3, 40, False)])
def test_native_step_loop_equals_stepwise(hs, every, nsteps, use_graph):
    """hs2_run_steps (native loop, CUDA-graph replay, device-side record counter) against one run_adi_steps call per
    step: same kernels on the same data -> identical bits, whatever the recording period (odd periods replay two
    periods per graph) and wherever steps with a source (taken one by one) cut the run into segments."""
    import torch
    prob = problems.sources_demo(hs)       # sources fire at steps 0, 1..3, 2, 3: the first native segment starts at 4
    P, S = hs.setup(*prob["setup_args"])
    ve, vol, dt, dz = prob["volumetric_elements"], prob["volumetric"], prob["dt"], prob["dz"]
    probes = [(0, 0, 0), (5, 4, 3), (11, 9, 13)]
    T = torch.from_numpy(np.array(prob["T0"])).cuda()
    Tn, rec = hs.run_adi_steps_n(P, S, prob["t0"], dt, T, ve, vol, nsteps, probes=probes, surface_dz=dz, every=every,
                                 use_graph=use_graph)
    cur, want_p, want_s = T.clone(), [], []
    for n in range(nsteps):
        cur = hs.run_adi_steps(P, S, prob["t0"] + n * dt, dt, cur, ve, vol)
        if (n + 1) % every == 0:
            want_p.append([float(cur[p]) for p in probes])
            # on numpy, like the reference (torch divides a tensor by a scalar as a multiplication by 1/scalar)
            want_s.append(hs.surface_temperature.insulating_z_min_surface_temperature(cur.cpu().numpy(), dz))
    assert rec["step"] == [(r + 1) * every for r in range(nsteps // every)]
    assert np.array_equal(Tn.cpu().numpy(), cur.cpu().numpy())
    assert np.array_equal(rec["probes"], np.array(want_p))
    assert np.array_equal(rec["surface"], np.array(want_s))
    # the caller's tensor is left alone, and a second call (cached graph, other record buffers) gives the same
    assert np.array_equal(T.cpu().numpy(), np.array(prob["T0"]))
    Tn2, rec2 = hs.run_adi_steps_n(P, S, prob["t0"], dt, T, ve, vol, nsteps, probes=probes, surface_dz=dz, every=every,
                                   use_graph=use_graph)
    assert np.array_equal(Tn2.cpu().numpy(), Tn.cpu().numpy()) and np.array_equal(rec2["surface"], rec["surface"])


def test_native_step_loop_source_free_numpy(hs):
    """no sources at all: the whole run is one hs2_run_steps call (graph units + remainder); numpy in, numpy out"""
    import adi_oracle
    prob = problems.uniform_slab(hs, shape=(24, 20, 32), nsteps=37)
    P, S = hs.setup(*prob["setup_args"])
    Tn, rec = hs.run_adi_steps_n(P, S, prob["t0"], prob["dt"], prob["T0"], prob["volumetric_elements"], prob["volumetric"],
                                 37, probes=[(3, 2, 1)], every=5)
    assert isinstance(Tn, np.ndarray) and rec["probes"].shape == (7, 1) and "surface" not in rec
    want = adi_oracle.run(prob, nsteps=37)
    assert util.relerr(Tn, want) <= 1e-11
    assert abs(rec["probes"][-1, 0] - adi_oracle.run(prob, nsteps=35)[3, 2, 1]) <= 1e-11 * np.abs(want).max()


@pytest.mark.parametrize("name,kwargs,world,nsteps", [
    ("steelonfoam", dict(nz=64, ny=40, nx=48), 2, 6),
    ("uniform_slab", dict(shape=(128, 48, 64)), 4, 4),
    ("uniform_slab", dict(shape=(64, 40, 50)), 8, 3),
    ("composite", dict(nz=64, ny=32, nx=32, ply=8), 2, 4),
    ("steelonwater", dict(nz=64, ny=40, nx=48), 2, 4),
    ("curved_plate", dict(nz=32, ny=20, nx=28), 2, 4),
])
def test_slab_kernels_on_one_gpu(hs, name, kwargs, world, nsteps):
    """The kernels of the multi-GPU z-slab path, slabs run one after the other
    on this GPU (tests/slab_seq.py): same field as the single plan (1e-13) and
    as the oracle (1e-12)."""
    import adi_oracle
    import slab_seq
    prob = problems.ALL[name](hs, **kwargs)
    got = slab_seq.run(hs, prob, world, nsteps)
    one = util.run_b200(hs, prob, nsteps=nsteps)
    assert util.relerr(got, one) <= 1e-13
    assert util.relerr(got, adi_oracle.run(prob, nsteps=nsteps)) <= 1e-12


@pytest.mark.parametrize("name,kwargs", [
    ("uniform_slab", dict(shape=(96, 128, 160))),
    ("steelonwater", dict(nz=96, ny=64, nx=128)),
    ("composite", dict(nz=64, ny=96, nx=128)),
])
def test_constant_bank_tables_are_bit_identical(hs, name, kwargs):
    """uniform-chunk fast path (factors as constant operands, chunk_core.cuh) against the
    table-load path (HS2_FLAG_NO_UTAB): same values, same operation order -> same bits"""
    import os
    prob = problems.ALL[name](hs, **kwargs)
    old = os.environ.get("HS2_UTAB_AXES")
    os.environ["HS2_UTAB_AXES"] = "xyz"
    try:
        a, plan = _run_plan(hs, prob, 3, flags=0)
        b, _ = _run_plan(hs, prob, 3, flags=4)        # HS2_FLAG_NO_UTAB
    finally:
        if old is None:
            del os.environ["HS2_UTAB_AXES"]
        else:
            os.environ["HS2_UTAB_AXES"] = old
    from heatsim2_b200 import _cabi
    assert all(int(plan.copy_table(a, _cabi.TAB_UCODE, np.uint8).sum()) > 0 for a in range(3)), "no uniform chunks found"
    assert np.array_equal(a, b)


@pytest.mark.parametrize("name,kwargs,nsteps", [
    ("curved_map", dict(), 6),                            # one curvature pair per (j, i): every cell its own equation
    ("steelonwater", dict(nz=16, ny=14, nx=48), 3),
    ("sources_demo", dict(), 6),
    ("uniform_slab", dict(shape=(9, 20, 31)), 3),
])
def test_four_byte_class_ids_whole_line_kernels(hs, monkeypatch, name, kwargs, nsteps):
    """More than 65536 equation classes (a curvature MAP on a large grid) run with 4-byte class ids on the whole-line
    kernels (VERDICT r01 'per-cell curvature').  Forced here on small problems (HS2_WIDE_IDS=1): same fields as the
    oracle, and for the smooth map the reference's own golden vector (tests/golden/c7_curved_map.npz)."""
    import adi_oracle
    monkeypatch.setenv("HS2_WIDE_IDS", "1")
    prob = problems.ALL[name](hs, **kwargs)
    got, plan = _run_plan(hs, prob, nsteps)
    assert plan.wide_ids and plan.class_id.element_size() == 4
    assert plan.last_kernels() == ("whole-line",) * 3
    assert util.relerr(got, adi_oracle.run(prob, nsteps=nsteps)) <= 1e-12 * nsteps
    if name == "curved_map":
        z, meta = util.load_golden("c7_curved_map")
        assert util.relerr(got, z["T_%d" % nsteps]) <= 1e-12 * nsteps


def test_curvature_map_class_count(hs):
    """the smooth map: one geometry class per (curvature pair, layer) - the two-byte ids of the tile kernels"""
    prob = problems.curved_map(hs)
    P, S = hs.setup(*prob["setup_args"])
    nz, ny, nx = prob["shape"]
    assert P.plan.n_classes >= ny * nx * (nz - 1) // 2 and not P.plan.wide_ids

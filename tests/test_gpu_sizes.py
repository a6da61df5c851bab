"""GPU parity at the line lengths and kernel variants the benchmarks run
("parity gaps"): 512- and 1024-row lines on every axis against the
oracle, the slab kernels at the SCALE run's chunking, and the SURVEY 8(d)
down-scaled configurations (C2 64^3 / 128^3 x 1000 steps, C3 128^3, C4
64x128x128 with per-step surface temperature) against digests produced by the
unmodified reference (tests/golden/big_*.npz, generator committed).

Every case asserts the chunking and the kernel variant that ran
(hs2_plan_last_kernel), so a silent change of code path fails the test."""
import numpy as np
import pytest

import problems
import util

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-12      # BASELINE.json: relative, per step
TOL_RUN = 1e-10       # after 1000 free-running steps


@pytest.fixture(scope="module")
def hs():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import heatsim2_b200
    from heatsim2_b200 import _cabi
    _cabi.lib()
    return heatsim2_b200


def _steps_vs_oracle(hs, prob, nsteps, restart=True):
    """run ``nsteps``; restart=True: every step starts from the oracle's state (<= 1e-12)"""
    import torch
    import adi_oracle
    O = adi_oracle.setup(*prob["setup_args"])
    P, S = hs.setup(*prob["setup_args"])
    T = np.array(prob["T0"])
    worst = 0.0
    for it in range(nsteps):
        t = prob["t0"] + prob["dt"] * it
        want = O.step(t, prob["dt"], T)
        got = hs.run_adi_steps(P, S, t, prob["dt"], torch.from_numpy(T).cuda(),
                               prob["volumetric_elements"], prob["volumetric"]).cpu().numpy()
        worst = max(worst, util.relerr(got, want))
        T = want if restart else got
    return worst, P.plan


# (problem, kwargs, long axis 0=x 1=y 2=z, expected (chunk rows, chunks) on it, expected kernel on it)
LONG_LINES = [
    ("uniform_slab", dict(shape=(8, 16, 512)), 0, (16, 32), "x-warp"),
    ("uniform_slab", dict(shape=(21, 47, 512)), 0, (16, 32), "x-warp"),     # odd plane count, ragged row patches, several patches per group
    ("steelonfoam", dict(nz=12, ny=20, nx=512, nsteps=3), 0, (16, 32), "x-warp"),   # delamination gap: lines with a class change inside
    ("uniform_slab", dict(shape=(8, 512, 16)), 1, (32, 16), "tile-tma"),
    ("uniform_slab", dict(shape=(512, 8, 16)), 2, (32, 16), "tile-cpasync"),
    ("uniform_slab", dict(shape=(8, 16, 1024)), 0, (16, 64), "x-warp"),       # two warps per line
    ("uniform_slab", dict(shape=(21, 47, 1024)), 0, (16, 64), "x-warp"),      # odd plane count, many patches per group
    ("uniform_slab", dict(shape=(8, 1024, 16)), 1, (32, 32), "tile-tma-512"),
    ("uniform_slab", dict(shape=(1024, 8, 16)), 2, (32, 32), "tile-cpasync-512"),
    # several unique lines along the long axis (material change, thin layer, delamination)
    ("steelonwater", dict(nz=8, ny=16, nx=512), 0, (16, 32), "x-warp"),
    ("steelonwater", dict(nz=9, ny=14, nx=512), 0, (16, 32), "x-warp"),      # odd plane count, ragged row patches
    ("composite", dict(nz=19, ny=10, nx=512), 0, (16, 32), "x-warp"),
    ("composite", dict(nz=19, ny=10, nx=256), 0, (16, 16), "x-warp"),          # two lines per warp
    ("uniform_slab", dict(shape=(21, 47, 256)), 0, (16, 16), "x-warp"),
    ("steelonfoam", dict(nz=12, ny=20, nx=256, nsteps=3), 0, (16, 16), "x-warp"),
    ("steelonwater", dict(nz=9, ny=14, nx=128), 0, (16, 8), "x-tma"),
    ("steelonwater", dict(nz=8, ny=512, nx=16), 1, (32, 16), "tile-tma"),
    ("composite", dict(nz=512, ny=8, nx=16), 2, (32, 16), "tile-cpasync"),
    ("composite", dict(nz=1024, ny=8, nx=16), 2, (32, 32), "tile-cpasync-512"),
    ("steelonwater", dict(nz=8, ny=16, nx=1024), 0, (16, 64), "x-warp"),       # lines that are not ghost-uniform: tables from global memory
    ("steelonwater", dict(nz=9, ny=14, nx=1024), 0, (16, 64), "x-warp"),
    ("composite", dict(nz=19, ny=10, nx=1024), 0, (16, 64), "x-warp"),
    ("steelonfoam", dict(nz=12, ny=20, nx=1024, nsteps=3), 0, (16, 64), "x-warp"),   # class change inside chunks, also across the middle
    # the benchmark's own tile shape on all three axes at once would be 512^3; 3 x (two long axes) instead
    ("uniform_slab", dict(shape=(4, 512, 512)), 1, (32, 16), "tile-tma"),
    ("uniform_slab", dict(shape=(512, 4, 512)), 2, (32, 16), "tile-cpasync"),
]


# the LSU-fed folded x kernel (kernels_xf.cu) stays the path of lines the TMA kernels do not take; forced here at 512
FOLD_LINES = [
    ("uniform_slab", dict(shape=(8, 16, 512)), 0, (32, 16), "x-fold"),
    ("steelonwater", dict(nz=8, ny=16, nx=512), 0, (32, 16), "x-fold"),
    ("uniform_slab", dict(shape=(8, 16, 1024)), 0, (32, 32), "x-fold"),
    ("steelonwater", dict(nz=8, ny=16, nx=1024), 0, (32, 32), "x-fold"),
]

# the TMA-fed patch kernel (kernels_xt.cu) runs the steps with a volumetric source and the lines of 16..496 cells;
# forced here at 512 (HS2_X_KERNEL=tma)
PATCH_LINES = [
    ("uniform_slab", dict(shape=(8, 16, 512)), 0, (16, 32), "x-tma"),
    ("steelonwater", dict(nz=9, ny=14, nx=512), 0, (16, 32), "x-tma"),
    ("composite", dict(nz=19, ny=10, nx=256), 0, (16, 16), "x-tma"),
]


@pytest.mark.parametrize("name,kwargs,axis,chunk,kernel", PATCH_LINES)
def test_patch_x_kernel_vs_oracle(hs, monkeypatch, name, kwargs, axis, chunk, kernel):
    monkeypatch.setenv("HS2_X_KERNEL", "tma")
    test_long_lines_vs_oracle(hs, name, kwargs, axis, chunk, kernel)


@pytest.mark.parametrize("shape", ["33", "52", "42", "24"])
def test_warp_x_kernel_shapes_vs_oracle(hs, shape):
    """both patch shapes of the warp-per-line kernel (HS2_XW_SHAPE is read once per process: subprocess)"""
    import os
    import subprocess
    import sys
    code = ("import sys; sys.path[:0] = %r; import numpy as np, torch, heatsim2_b200 as hs, adi_oracle, problems, util\n"
            "for name, kw in (('uniform_slab', dict(shape=(21, 47, 512))), ('steelonwater', dict(nz=9, ny=14, nx=512)),\n"
            "                 ('steelonfoam', dict(nz=12, ny=20, nx=512, nsteps=3)), ('composite', dict(nz=19, ny=10, nx=512)),\n"
            "                 ('uniform_slab', dict(shape=(2, 5, 512))), ('uniform_slab', dict(shape=(1, 9, 512)))):\n"
            "    prob = problems.ALL[name](hs, **kw)\n"
            "    got = util.run_b200(hs, prob, nsteps=3)\n"
            "    err = util.relerr(got, adi_oracle.run(prob, nsteps=3))\n"
            "    assert err <= 3e-12, (name, err)\n"
            "import slab_seq\n"
            "prob = problems.ALL['steelonwater'](hs, nz=16, ny=14, nx=512)\n"       # interior / boundary plane ranges (slabs)
            "err = util.relerr(slab_seq.run(hs, prob, 2, 3), adi_oracle.run(prob, nsteps=3))\n"
            "assert err <= 3e-12, ('slabs', err)\n"
            "print('ok')\n") % ([p for p in sys.path if p],)
    env = dict(os.environ, HS2_XW_SHAPE=shape)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("name,kwargs,axis,chunk,kernel", FOLD_LINES)
def test_folded_x_kernel_vs_oracle(hs, monkeypatch, name, kwargs, axis, chunk, kernel):
    monkeypatch.setenv("HS2_X_KERNEL", "fold")
    test_long_lines_vs_oracle(hs, name, kwargs, axis, chunk, kernel)


@pytest.mark.parametrize("name,kwargs,axis,chunk,kernel", LONG_LINES)
def test_long_lines_vs_oracle(hs, name, kwargs, axis, chunk, kernel):
    prob = problems.ALL[name](hs, **kwargs)
    err, plan = _steps_vs_oracle(hs, prob, 3)
    assert plan.chunk[axis] == chunk, plan.chunk
    assert plan.last_kernels()[axis] == kernel, plan.last_kernels()
    if name != "uniform_slab":
        assert plan.n_unique[axis] >= 2, plan.n_unique
    assert err <= TOL_STEP, err


@pytest.mark.parametrize("name,kwargs,world,slab_chunk,nsteps", [
    ("uniform_slab", dict(shape=(64, 16, 1024)), 8, 8, 3),        # VERDICT: x lines of 1024 rows inside slabs
    ("uniform_slab", dict(shape=(1024, 8, 16)), 8, 32, 3),        # SCALE N=8: 128-plane slabs, 32 global chunks of 32
    ("uniform_slab", dict(shape=(1024, 8, 16)), 4, 32, 3),        # SCALE N=4: 256-plane slabs
    ("uniform_slab", dict(shape=(1024, 8, 16)), 2, 32, 3),        # SCALE N=2: 512-plane slabs
    ("composite", dict(nz=512, ny=8, nx=16), 8, 32, 3),
])
def test_slab_kernels_long_lines(hs, name, kwargs, world, slab_chunk, nsteps):
    """slab path (hs2_sweep_x_part + hs2_sweep_z_forward/backward) at the SCALE bench's chunking"""
    import adi_oracle
    import slab_seq
    from heatsim2_b200.plan import slab_chunk as choose
    prob = problems.ALL[name](hs, **kwargs)
    assert choose(prob["shape"][0] // world) == slab_chunk
    got = slab_seq.run(hs, prob, world, nsteps)
    assert util.relerr(got, adi_oracle.run(prob, nsteps=nsteps)) <= TOL_STEP * nsteps


@pytest.mark.parametrize("case", util.big_golden_cases())
def test_survey_sizes_vs_reference_digests(hs, case):
    """free-running runs at the down-scaled BASELINE sizes against the
    unmodified reference (digest: sub-sampled field, sum, 4 random projections)."""
    import torch
    z, meta = util.load_golden(case)
    prob = problems.ALL[meta["problem"]](hs, **meta["kwargs"])
    P, S = hs.setup(*prob["setup_args"])
    T = torch.from_numpy(np.array(prob["T0"])).cuda()
    nxt = torch.empty_like(T)
    cols = meta.get("surface_cols")
    hist = []
    for it in range(max(meta["steps"])):
        hs.run_adi_steps(P, S, prob["t0"] + prob["dt"] * it, prob["dt"], T, prob["volumetric_elements"],
                         prob["volumetric"], out=nxt)
        T, nxt = nxt, T
        if cols:
            # C4: surface temperature evaluated on the device tensor every step, by the library's
            # observation kernel (hs2_observe), which must give the bits of the reference's formula
            surf = torch.empty(T.shape[1:], dtype=T.dtype, device=T.device)
            P.plan.observe(T, None, None, surf, prob["dz"])
            if it % 50 == 0:
                assert np.array_equal(surf.cpu().numpy(), hs.surface_temperature.insulating_z_min_surface_temperature(
                    T.cpu().numpy(), prob["dz"]))      # numpy, like the reference: true division
            hist.append(torch.stack([surf[tuple(c)] for c in cols]))
        if (it + 1) in meta["steps"]:
            tol = TOL_STEP if it == 0 else TOL_RUN
            util.check_digest(T.cpu().numpy(), z, it + 1, meta["sub"], tol)
    if cols:
        got = torch.stack(hist).cpu().numpy()
        assert util.relerr(got, z["surface_hist"]) <= TOL_RUN
    assert all(k.startswith("tile") or k in ("x-tma", "x-warp") for k in P.plan.last_kernels()), P.plan.last_kernels()


def test_c2_64_1000_steps_vs_oracle(hs):
    """the same C2 recipe at 64^3 for 1000 free-running steps against the numpy oracle run here"""
    import adi_oracle
    prob = problems.uniform_slab(hs, n=64, nsteps=1000)
    got = util.run_b200(hs, prob)
    assert util.relerr(got, adi_oracle.run(prob)) <= TOL_RUN


def test_more_than_eight_active_source_classes(hs):
    """9+ volumetric regions active in one step (the reference allows 256)"""
    import adi_oracle
    prob = problems.uniform_slab(hs, shape=(12, 20, 24))
    args = list(prob["setup_args"])
    nsrc = 12
    volumetric = ((hs.NO_SOURCE,),) + tuple((hs.STEPPED_SOURCE, 0.0, 1.0, 1e6 * (s + 1)) for s in range(nsrc))
    ve = np.zeros(prob["shape"], dtype=np.uint8)
    for s in range(nsrc):
        ve[s % 12, :, 2 * s:2 * s + 2] = s + 1
    args[12], args[17] = volumetric, ve
    prob = dict(prob, setup_args=tuple(args), volumetric=volumetric, volumetric_elements=ve)
    err, plan = _steps_vs_oracle(hs, prob, 2)
    assert err <= TOL_STEP


def test_out_argument_is_validated(hs):
    import torch
    prob = problems.uniform_slab(hs, shape=(8, 16, 32))
    P, S = hs.setup(*prob["setup_args"])
    T = torch.zeros(prob["shape"], dtype=torch.float64, device="cuda")
    for bad in (torch.zeros(prob["shape"], dtype=torch.float32, device="cuda"),
                torch.zeros((8, 16, 64), dtype=torch.float64, device="cuda")[:, :, ::2],
                torch.zeros((8, 16, 31), dtype=torch.float64, device="cuda"),
                torch.zeros(prob["shape"], dtype=torch.float64)):
        with pytest.raises(ValueError):
            hs.run_adi_steps(P, S, 0.0, prob["dt"], T, prob["volumetric_elements"], prob["volumetric"], out=bad)

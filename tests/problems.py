"""Problem definitions shared by the parity tests, the golden-vector generator
and bench.py.  Every builder takes the package to build *for* (``hs`` = the
reference ``heatsim2`` from oracle/_ref, or ``heatsim2_b200``) because boundary
plug-ins are module objects of that package, and returns a dict with the
positional arguments of ``hs.setup`` plus run parameters.

Configurations follow BASELINE.json / SURVEY.md 8(d):
  C1 steelonfoam      demos/steelonfoam.py as shipped (80x40x48)
  C2 uniform_slab     one material, insulated box, random T0 + flash
  C3 steelonwater     demos/steelonwater.py geometry + thin insulating layer
  C4 composite        alternating anisotropic plies + delamination
"""
import numpy as np

STEEL = (40.0, 7.75e3, 466.0)
FOAM = (0.4, 40.0, 1500.0)
WATER = (20.0, 1e3, 418.0)      # numbers of demos/steelonwater.py:44-49 (as shipped)


def _insulate_outer(bz, by, bx, b=1):
    bx[:, :, 0] = b
    bx[:, :, -1] = b
    by[:, 0, :] = b
    by[:, -1, :] = b
    bz[0, :, :] = b
    bz[-1, :, :] = b


def _pack(hs, grid, dt, materials, boundaries, volumetric, elems, T0, nsteps, t0=0.0, **extra):
    (dz, dy, dx, z, y, x) = grid
    me, bz, by, bx, ve = elems
    d = dict(setup_args=(z[0], y[0], x[0], dz, dy, dx, len(z), len(y), len(x), dt,
                         materials, boundaries, volumetric, me, bz, by, bx, ve),
             shape=(len(z), len(y), len(x)), dz=dz, dy=dy, dx=dx, dt=dt, t0=t0,
             materials=materials, boundaries=boundaries, volumetric=volumetric,
             material_elements=me, volumetric_elements=ve, T0=T0, nsteps=nsteps)
    d.update(extra)
    return d


def steelonfoam(hs, nz=80, ny=40, nx=48, nsteps=999):
    """demos/steelonfoam.py:21-113."""
    dt = 0.01
    (dz, dy, dx, z, y, x, zgrid, ygrid, xgrid, z_bnd, y_bnd, x_bnd) = hs.build_grid(
        0, 10.8e-3, nz, -0.05, 0.05, ny, -.06, .06, nx)[:12]
    ygrid2d, xgrid2d = np.meshgrid(y, x, indexing="ij")
    fi = int(np.argmin(np.abs(3.175e-3 - z_bnd)))
    materials = ((hs.TEMPERATURE_COMPUTE,) + STEEL, (hs.TEMPERATURE_COMPUTE,) + FOAM, (hs.TEMPERATURE_FIXED,))
    boundaries = ((hs.boundary_conducting,), (hs.boundary_insulating,))
    volumetric = ((hs.NO_SOURCE,), (hs.IMPULSE_SOURCE, 0.0, 10e3 / dz))
    me, bz, by, bx, ve = hs.zero_elements(nz, ny, nx)
    me[fi:, :, :] = 1
    _insulate_outer(bz, by, bx)
    ve[0, :, :] = 1
    bz[fi, :, :][(ygrid2d > 0) & (ygrid2d < 10e-3) & (xgrid2d > 0) & (xgrid2d < 10e-3)] = 1
    return _pack(hs, (dz, dy, dx, z, y, x), dt, materials, boundaries, volumetric, (me, bz, by, bx, ve),
                 np.zeros((nz, ny, nx)), nsteps, probes=((0, 22, 25), (0, 37, 43)))


def uniform_slab(hs, n=64, nsteps=10, seed=1234, shape=None, random_T0=True):
    """steel cube, d=1e-4, insulated, dt=0.01, random T0 +
    flash of 10 kJ/m^2 on layer 0 at t=0."""
    nz, ny, nx = shape if shape is not None else (n, n, n)
    d = 1e-4
    dt = 0.01
    g = hs.build_grid_min_step_edge(0, d, nz, 0, d, ny, 0, d, nx)
    z, y, x = g[0], g[1], g[2]
    materials = ((hs.TEMPERATURE_COMPUTE,) + STEEL,)
    boundaries = ((hs.boundary_conducting,), (hs.boundary_insulating,))
    volumetric = ((hs.NO_SOURCE,), (hs.IMPULSE_SOURCE, 0.0, 10e3 / d))
    me, bz, by, bx, ve = hs.zero_elements(nz, ny, nx)
    _insulate_outer(bz, by, bx)
    ve[0, :, :] = 1
    T0 = np.random.default_rng(seed).random((nz, ny, nx)) if random_T0 else np.zeros((nz, ny, nx))
    return _pack(hs, (d, d, d, z, y, x), dt, materials, boundaries, volumetric, (me, bz, by, bx, ve), T0, nsteps)


def steelonwater(hs, nz=80, ny=20, nx=24, nsteps=20, h=1000.0, seed=7):
    """demos/steelonwater.py:33-111 geometry (steel block,
    water pocket, FIXED last layer, insulated faces, dt=0.05) plus boundary
    class 2 = thin insulating layer ``h`` on every steel/water face."""
    dt = 0.05
    (dz, dy, dx, z, y, x, zgrid, ygrid, xgrid, z_bnd, y_bnd, x_bnd) = hs.build_grid(
        0, 50.8e-3, nz, -0.1, 0.1, ny, -.12, .12, nx)[:12]
    materials = ((hs.TEMPERATURE_COMPUTE,) + STEEL, (hs.TEMPERATURE_COMPUTE,) + WATER, (hs.TEMPERATURE_FIXED,))
    boundaries = ((hs.boundary_conducting,), (hs.boundary_insulating,), (hs.boundary_thininsulatinglayer, h))
    volumetric = ((hs.NO_SOURCE,), (hs.IMPULSE_SOURCE, 0.0, 10e3 / dz))
    me, bz, by, bx, ve = hs.zero_elements(nz, ny, nx)
    water = (xgrid > 0) & (xgrid < 10e-3) & (ygrid > 0) & (ygrid < 10e-3) & (zgrid > 9.525e-3)
    me[water] = 1
    me[-1, :, :] = 2
    # thin layer wherever steel meets water
    w = (me == 1)
    s = (me == 0)
    bz[1:-1][(w[1:] & s[:-1]) | (s[1:] & w[:-1])] = 2
    by[:, 1:-1][(w[:, 1:] & s[:, :-1]) | (s[:, 1:] & w[:, :-1])] = 2
    bx[:, :, 1:-1][(w[:, :, 1:] & s[:, :, :-1]) | (s[:, :, 1:] & w[:, :, :-1])] = 2
    _insulate_outer(bz, by, bx)
    ve[0, :, :] = 1
    T0 = np.random.default_rng(seed).random((nz, ny, nx))
    return _pack(hs, (dz, dy, dx, z, y, x), dt, materials, boundaries, volumetric, (me, bz, by, bx, ve), T0, nsteps)


def composite(hs, nz=32, ny=64, nx=64, nsteps=20, ply=8, seed=11):
    """plies A/B of axis-aligned tensors alternating every
    ``ply`` z-layers (numbers of demos/ktest.py), conducting_anisotropic
    interior, insulated faces + a rectangular delamination on a ply
    interface, flash on layer 0."""
    dt = 0.01
    dz = 0.125e-3
    dy = dx = 0.5e-3
    g = hs.build_grid_min_step_edge(0, dz, nz, 0, dy, ny, 0, dx, nx)
    z, y, x = g[0], g[1], g[2]
    rho, c = 1.75e3, 730.0
    KA = np.diag((0.71, 0.71, 5.1))
    KB = np.diag((0.71, 5.1, 0.71))
    materials = ((hs.TEMPERATURE_COMPUTE, KA, rho, c), (hs.TEMPERATURE_COMPUTE, KB, rho, c))
    boundaries = ((hs.boundary_conducting_anisotropic,), (hs.boundary_insulating,))
    volumetric = ((hs.NO_SOURCE,), (hs.IMPULSE_SOURCE, 0.0, 10e3 / dz))
    me, bz, by, bx, ve = hs.zero_elements(nz, ny, nx)
    me[((np.arange(nz) // ply) % 2 == 1), :, :] = 1
    _insulate_outer(bz, by, bx)
    kd = min(nz - 1, 2 * ply)
    bz[kd, ny // 4: ny // 2, nx // 4: (3 * nx) // 4] = 1
    ve[0, :, :] = 1
    T0 = np.random.default_rng(seed).random((nz, ny, nx))
    return _pack(hs, (dz, dy, dx, z, y, x), dt, materials, boundaries, volumetric, (me, bz, by, bx, ve), T0, nsteps)


def sources_demo(hs, nz=12, ny=10, nx=14, seed=3):
    """All four volumetric source kinds on a small two-material block with a
    FIXED sink (covers alternatingdirection_c_pyx.pyx:294-383)."""
    dt = 0.02
    (dz, dy, dx, z, y, x, zgrid, ygrid, xgrid) = hs.build_grid(0, 6e-3, nz, -5e-3, 5e-3, ny, -7e-3, 7e-3, nx)[:9]
    materials = ((hs.TEMPERATURE_COMPUTE,) + STEEL, (hs.TEMPERATURE_COMPUTE,) + FOAM, (hs.TEMPERATURE_FIXED,))
    boundaries = ((hs.boundary_conducting,), (hs.boundary_insulating,))
    volumetric = ((hs.NO_SOURCE,),
                  (hs.IMPULSE_SOURCE, 0.0, 10e3 / dz),
                  (hs.STEPPED_SOURCE, 0.02 - dt / 2, 0.08 - dt / 2, 2e7),
                  (hs.IMPULSE_POINT_SOURCE_JOULES, 0.04, 1e-3),
                  (hs.SPATIALLY_Z_DECAYING_TEMPORAL_IMPULSE, 0.06, np.array([1.0, 0.0, 0.0]), 0.0, zgrid, dz, 5e3, 1.5e-3))
    me, bz, by, bx, ve = hs.zero_elements(nz, ny, nx)
    me[nz // 2:, :, :] = 1
    me[-1, :, :] = 2
    _insulate_outer(bz, by, bx)
    ve[0, :, :nx // 2] = 1
    ve[1:3, :, nx // 2:] = 2
    ve[4, 5, 6] = 3
    ve[:, :2, :] = 4
    T0 = np.random.default_rng(seed).random((nz, ny, nx))
    return _pack(hs, (dz, dy, dx, z, y, x), dt, materials, boundaries, volumetric, (me, bz, by, bx, ve), T0, 6)


def curved_plate(hs, nz=24, ny=20, nx=28, nsteps=12, seed=21):
    """Curved-surface mode (crank_nicolson.pyx:388-458): steel on foam under a
    top surface with curvature 1/(50 mm) about x everywhere and two different
    curvatures about y (concave left half, convex right half), so the cell
    geometry - and with it the equation class - changes with depth.  Flash on
    layer 0 at t=0, a point source (Joules; divides by the per-cell volume) at
    t=2*dt, an insulating gap on the interface."""
    dt = 0.01
    (dz, dy, dx, z, y, x, zgrid, ygrid, xgrid, z_bnd, y_bnd, x_bnd) = hs.build_grid(
        0, 10.8e-3, nz, -0.05, 0.05, ny, -.06, .06, nx)[:12]
    fi = nz // 3
    materials = ((hs.TEMPERATURE_COMPUTE,) + STEEL, (hs.TEMPERATURE_COMPUTE,) + FOAM)
    boundaries = ((hs.boundary_conducting,), (hs.boundary_insulating,))
    volumetric = ((hs.NO_SOURCE,), (hs.IMPULSE_SOURCE, 0.0, 10e3 / dz), (hs.IMPULSE_POINT_SOURCE_JOULES, 2 * dt, 2e-3))
    me, bz, by, bx, ve = hs.zero_elements(nz, ny, nx)
    me[fi:, :, :] = 1
    _insulate_outer(bz, by, bx)
    bz[fi, ny // 4: ny // 2, nx // 4: nx // 2] = 1
    ve[0, :, :] = 1
    ve[2, ny // 2, nx // 3] = 2
    cy = np.full((ny, nx), 1.0 / 50e-3)
    cx = np.where(np.arange(nx)[None, :] < nx // 2, 1.0 / 80e-3, -1.0 / 200e-3) * np.ones((ny, nx))
    T0 = np.random.default_rng(seed).random((nz, ny, nx))
    d = _pack(hs, (dz, dy, dx, z, y, x), dt, materials, boundaries, volumetric, (me, bz, by, bx, ve), T0, nsteps)
    d["setup_args"] = d["setup_args"] + (cy, cx)
    return d


def curved_map(hs, nz=8, ny=10, nx=12, nsteps=6, seed=22):
    """Curved-surface mode with a SMOOTH curvature map: every (j, i) has its own pair of curvatures (a measured
    surface), so every cell has its own geometry and - with the depth - its own equation
    (crank_nicolson.pyx:388-458 evaluates it per cell).  Otherwise the curved_plate recipe on a small grid."""
    d = curved_plate(hs, nz=nz, ny=ny, nx=nx, nsteps=nsteps, seed=seed)
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    cy = (1.0 / 50e-3) * (1.0 + 0.25 * np.sin(0.7 * jj + 0.31 * ii) + 0.01 * jj)
    cx = (1.0 / 120e-3) * np.cos(0.45 * ii - 0.2 * jj) - 0.002 * ii * jj
    d["setup_args"] = d["setup_args"][:-2] + (cy, cx)
    return d


ALL = {"curved_plate": curved_plate, "curved_map": curved_map, "steelonfoam": steelonfoam, "uniform_slab": uniform_slab, "steelonwater": steelonwater,
       "composite": composite, "sources_demo": sources_demo}

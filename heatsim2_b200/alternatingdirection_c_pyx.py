"""ADI stage bookkeeping and the time-step entry point.

Host-side mirror of the reference's ``heatsim2/alternatingdirection_c_pyx.pyx``
(+ the C container ``alternatingdirection_c.c``) with the same public names:

    adi_params :75      pyadi_step :120     adi_setup :418
    adi_expressions :475                    add_equation_to_adi_matrices :447
    run_adi_steps :287

What changed underneath: the reference stores, per stage, an ``n x 3``
tridiagonal, COO/CSR matrices B, C0, C1 and a vector D (1.5 kB per cell) and
each step does 3 x {SpMV + one length-n Thomas chain} on one CPU thread.  Here a
stage is described by per-equation-class coefficients (a few dozen rows of 8
doubles) plus per-unique-line Thomas factors; ``run_adi_steps`` hands the field
to the CUDA library (``_cabi`` -> libhs2b200.so).  The matrices the reference
exposes as attributes (``Amat``, ``Bmat``, ``Cmats``, ``Dvec``, ``Lmat``,
``Umat``) can still be materialised on the host for inspection of small grids.
"""
import numpy as np

from . import expression
from .expression import crank_subst_in_groups, eliminate_groups

# sweep order of the three Douglas stages: (axis name, permuteorder)
_STAGES = (("x", (0, 1, 2)), ("y", (2, 0, 1)), ("z", (1, 2, 0)))


class adi_params(object):
    """Problem-wide parameters shared by the stages (reference :75-85).
    Unknown keyword -> ValueError, as in the reference."""
    shape = None
    volume_array = None
    plan = None          # heatsim2_b200.plan.AdiPlan, attached by setup()

    def __init__(self, **kwargs):
        for key in kwargs:
            if not hasattr(self, key):
                raise ValueError("Unknown parameter %s (must add to class definition)" % key)
            setattr(self, key, kwargs[key])


class pyadi_step(object):
    """One ADI stage (reference :120-284).  Keeps the reference's descriptive
    attributes; the numerical content lives in ``ADI_params.plan``."""

    def __init__(self, ADI_params, permuteorder, stepnum):
        self.ADI_params = ADI_params
        self.stepnum = stepnum
        self.permuteorder = tuple(permuteorder)
        self.permutedshape = [ADI_params.shape[a] for a in permuteorder]
        self.invpermuteorder = tuple(int(a) for a in np.argsort(permuteorder))

    def add_equation(self, posindex, eqdict):
        """Reference method (:177-209): the equation ``{variable name: coefficient}`` of THIS stage for cell
        ``posindex = (k, j, i)``.  The reference hands it to C ``add_equation``, which appends COO triplets; here it
        is recorded, and once a cell has its equation for all three stages the cell joins an equation class of the
        plan builder (same rules, ``parse_stage_dict``).  ``finalize()`` of any stage then compiles the plan."""
        builder = getattr(self.ADI_params, "_cell_builder", None)
        if builder is None:
            from .plan import CellwiseBuilder
            builder = self.ADI_params._cell_builder = CellwiseBuilder(self.ADI_params.shape)
        builder.add_stage(self.stepnum, tuple(int(v) for v in posindex), dict(eqdict))

    def finalize(self):
        """The reference converts COO->CSR and LU-factors here (:212-282).
        Plans made by setup() are complete already; when equations were added
        cell by cell (add_equation_to_adi_matrices) the first finalize() turns
        them into the class tables."""
        builder = getattr(self.ADI_params, "_cell_builder", None)
        if builder is not None and self.ADI_params.plan is None:
            self.ADI_params.plan = builder.finish(None, self.ADI_params.volume_array)
        return None

    # --- host-side materialisation of the reference's matrices (inspection /
    #     tests on small grids only; never used by run_adi_steps) -------------
    def _mats(self):
        plan = self.ADI_params.plan
        if plan is None:
            raise RuntimeError("stage has no plan: call setup() first")
        return plan.reference_matrices(self.stepnum)

    @property
    def Amat(self):
        return self._mats()["A"]

    @property
    def Bmat(self):
        return self._mats()["B"]

    @property
    def Cmats(self):
        return self._mats()["C"]

    @property
    def Dvec(self):
        return self._mats()["D"]

    @property
    def Lmat(self):
        from . import tridiag
        return tridiag.tridiaglu_host(self.Amat)[0]

    @property
    def Umat(self):
        from . import tridiag
        return tridiag.tridiaglu_host(self.Amat)[1]


def adi_setup(shape, volume_array):
    """Three stages, implicit in x, then y, then z (Douglas eq. 3.1a-c;
    reference :418-445)."""
    ADI_params = adi_params(shape=tuple(int(s) for s in shape), volume_array=volume_array)
    ADI_steps = [pyadi_step(ADI_params, perm, s) for s, (_, perm) in enumerate(_STAGES)]
    return (ADI_params, ADI_steps)


def adi_expressions(spatial_expression, time_expression, unaligned_anisotropic=False):
    """Stage equations from one cell's heat balance (reference :475-572).

    Stage s time-averages (Crank-Nicolson) the flux groups of sweep directions
    0..s and solves for ``T555p<s>``."""
    if unaligned_anisotropic:
        # The reference's 15-stage branch (:511-564) cannot run: adi_setup()
        # creates 3 stages and add_equation_to_adi_matrices asserts equal
        # counts (:452).  There is no behaviour to be compatible with.
        raise NotImplementedError("unaligned anisotropic conduction is not supported "
                                  "(the reference's 15-stage path is non-functional)")
    spatial, times = [], []
    cur = spatial_expression
    for s, members in enumerate((("T556", "T555", "T554"), ("T565", "T555", "T545"), ("T655", "T555", "T455"))):
        cur = crank_subst_in_groups(cur, members, s)
        spatial.append(cur)
        times.append(expression.subst(time_expression, "T555p", "T555p%d" % s))
    spatial[-1] = spatial[-1].fullreduce()
    # tensor conductivities not aligned with the grid leave cross-derivative groups behind
    assert expression.no_groups(spatial[-1])
    return (tuple(spatial), tuple(times))


def stage_dicts(spatial_expressions, time_expressions):
    """``{variable: coefficient}`` of every stage (what the reference caches in
    ``aetam_cache``, :454-465)."""
    return [eliminate_groups(sp + tm).dictform() for sp, tm in zip(spatial_expressions, time_expressions)]


_OFFS = {"4": -1, "5": 0, "6": 1}


def parse_stage_dict(eqdict, stepnum):
    """Sort one stage dictionary into matrix entries by the rules of the
    reference's C ``add_equation`` (alternatingdirection_c.c:102-199).

    Returns ``(A, B, C, D)``: ``A`` maps the offset along the sweep axis to the
    tridiagonal entry (sign flipped, :147); ``B`` and ``C[sol]`` map a
    ``(dz,dy,dx)`` offset to the summed coefficient; ``D`` is the source weight."""
    sweep_axis = _STAGES[stepnum][1][2]            # 2 (x), 1 (y), 0 (z) in (z,y,x) digit order
    A, B, C, D = {}, {}, {}, 0.0
    for name, value in eqdict.items():
        if name == "volumetric_source":
            D = value
        elif name == "":
            if value != 0.0:
                raise ValueError("equation has a constant term %g; boundaries must be homogeneous" % value)
        else:
            if name[0] != "T" or len(name) < 4 or any(ch not in _OFFS for ch in name[1:4]):
                raise ValueError("unexpected variable %r in a stage equation" % name)
            off = tuple(_OFFS[ch] for ch in name[1:4])
            tshift = name[4:5]
            sol = int(name[5:]) if tshift == "p" and len(name) > 5 else 0
            if tshift == "p" and sol == stepnum:
                if any(off[a] != 0 for a in range(3) if a != sweep_axis):
                    raise ValueError("implicit variable %r is not on the sweep axis of stage %d" % (name, stepnum))
                A[off[sweep_axis]] = -value
            elif tshift in ("m", ""):
                B[off] = B.get(off, 0.0) + value
            elif tshift == "p":
                if sol > stepnum:
                    raise ValueError("stage %d refers to later stage %d" % (stepnum, sol))
                C.setdefault(sol, {})
                C[sol][off] = C[sol].get(off, 0.0) + value
            else:
                raise ValueError("unexpected time suffix in %r" % name)
    return A, B, C, D


def class_coefficients(eqdicts, rtol=1e-9):
    """Reduce the three stage dictionaries of one equation class to
    ``(M, (gx-,gx+,gy-,gy+,gz-,gz+), D)`` and verify that the stages have the
    conservative 7-point Douglas structure the kernels implement."""
    parsed = [parse_stage_dict(d, s) for s, d in enumerate(eqdicts)]
    g = []
    M = None
    for s in range(3):
        A = parsed[s][0]
        lo, hi, dg = -2.0 * A.get(-1, 0.0), -2.0 * A.get(1, 0.0), A.get(0, 0.0)
        g += [lo, hi]
        Ms = dg - 0.5 * (lo + hi)
        M = Ms if M is None else M
        scale = abs(dg) + abs(lo) + abs(hi)
        if abs(Ms - M) > rtol * scale:
            raise NotImplementedError("stage diagonals disagree on the capacity term (%g vs %g)" % (Ms, M))
    D = parsed[0][3]
    gx, gy, gz = (g[0], g[1]), (g[2], g[3]), (g[4], g[5])
    sx, sy, sz = sum(gx), sum(gy), sum(gz)
    scale = abs(M) + abs(sx) + abs(sy) + abs(sz)

    def offs(axis, val):
        o = [0, 0, 0]
        o[axis] = val
        return tuple(o)

    def stencil(cx, cy, cz, centre):
        st = {(0, 0, 0): centre}
        for (axis, gg, c) in ((2, gx, cx), (1, gy, cy), (0, gz, cz)):
            for sgn, gv in zip((-1, 1), gg):
                if c * gv != 0.0:
                    st[offs(axis, sgn)] = c * gv
        return st

    expect = [
        (stencil(.5, 1, 1, M - .5 * sx - sy - sz), {}),
        (stencil(.5, .5, 1, M - .5 * sx - .5 * sy - sz), {0: stencil(.5, 0, 0, -.5 * sx)}),
        (stencil(.5, .5, .5, M - .5 * (sx + sy + sz)), {0: stencil(.5, 0, 0, -.5 * sx), 1: stencil(0, .5, 0, -.5 * sy)}),
    ]

    def same(got, want):
        for key in set(got) | set(want):
            if abs(got.get(key, 0.0) - want.get(key, 0.0)) > rtol * scale:
                return False
        return True

    for s in range(3):
        _, B, C, Ds = parsed[s]
        ok = same(B, expect[s][0]) and all(same(C.get(c, {}), expect[s][1].get(c, {})) for c in set(C) | set(expect[s][1]))
        if not ok or abs(Ds - D) > rtol * max(abs(D), 1.0):
            raise NotImplementedError(
                "stage %d of an equation class is not of the conservative 7-point Douglas form "
                "supported by the CUDA kernels: B=%r C=%r" % (s, B, C))
    return M, tuple(g), D


def add_equation_to_adi_matrices(ADI_params, ADI_steps, k, j, i, key_params, aetam_cache,
                                 spatial_expressions, time_expressions):
    """Reference signature (:447-472).  The vectorised setup() never calls this
    per cell; it is kept for code that assembles cell by cell: the equation is
    recorded in the plan builder attached to ``ADI_params``."""
    assert len(ADI_steps) == len(spatial_expressions)
    try:
        eqdicts = aetam_cache[key_params]
    except KeyError:
        eqdicts = stage_dicts(spatial_expressions, time_expressions)
        aetam_cache[key_params] = eqdicts
    builder = getattr(ADI_params, "_cell_builder", None)
    if builder is None:
        from .plan import CellwiseBuilder
        builder = ADI_params._cell_builder = CellwiseBuilder(ADI_params.shape)
    builder.add(k, j, i, key_params, eqdicts)


def run_adi_steps(ADI_params, ADI_steps, t, dt, Tarray, volumetric_elements, volumetric, out=None):
    """Advance ``Tarray`` (indexed [z,y,x], float64) by one ADI time step.

    Same call as the reference (:287-416).  ``Tarray`` may be a numpy array -
    it is copied to the device, stepped, and a new numpy array is returned,
    like the reference - or a CUDA ``torch.Tensor``, in which case the result
    is a new CUDA tensor and nothing crosses PCIe.  A host ``torch.Tensor``
    (pinned for full PCIe speed) is copied in and a host tensor comes back.
    ``out`` (extension) receives the result instead of a fresh allocation; for
    device tensors it may be ``Tarray`` itself."""
    plan = ADI_params.plan
    if plan is None:
        raise RuntimeError("ADI_params carries no plan; it must come from heatsim2_b200.setup()")
    return plan.run_step(t, dt, Tarray, volumetric_elements, volumetric, out=out)


def run_adi_steps_n(ADI_params, ADI_steps, t0, dt, Tarray, volumetric_elements, volumetric, nsteps,
                    probes=None, surface_dz=None, every=1, use_graph=True):
    """``nsteps`` time steps with the field resident on the device and
    observation on the device (extension; the reference's demos copy the whole
    field to the host every step and index it there, demos/steelonfoam.py:132-143).

    Step n (0-based) is taken at time ``t0 + n*dt``, like the demos' loops.
    ``probes``: list of (k, j, i) cells whose temperature is recorded after
    every ``every``-th step; ``surface_dz``: when given, the insulated z-min
    surface temperature (``surface_temperature.insulating_z_min_surface_temperature``)
    is evaluated on the device at the same instants.  Only the recorded values
    cross PCIe, once, at the end.

    Runs of steps without an active volumetric source go to the library in ONE
    call (``hs2_run_steps``: native loop, one observation kernel per recorded
    instant, replayed as a CUDA graph unless ``use_graph`` is false); steps with
    a source are taken one by one.

    Returns ``(T_final, record)`` with ``record = {"step": [...], "probes":
    ndarray [n_rec, n_probes], "surface": ndarray [n_rec, ny, nx]}`` (keys present
    when requested).  ``Tarray`` may be numpy (copied in once; result numpy) or a
    CUDA tensor (result CUDA tensor)."""
    import torch
    plan = ADI_params.plan
    if plan is None:
        raise RuntimeError("ADI_params carries no plan; it must come from heatsim2_b200.setup()")
    nsteps, every = int(nsteps), int(every)
    if every < 1:
        raise ValueError("every must be >= 1")
    if not hasattr(plan, "run_steps_device"):
        return _run_adi_steps_n_stepwise(ADI_params, ADI_steps, t0, dt, Tarray, volumetric_elements, volumetric, nsteps,
                                         probes, surface_dz, every)
    was_numpy = not isinstance(Tarray, torch.Tensor)
    plan.ensure_device(None if was_numpy or not Tarray.is_cuda else Tarray.device)
    with torch.cuda.device(plan._dev):
        if was_numpy:
            cur = torch.empty(plan.shape, dtype=torch.float64, device=plan._dev)
            plan.upload(Tarray, cur)
        else:
            if Tarray.dtype != torch.float64 or tuple(Tarray.shape) != plan.shape:
                raise ValueError("Tarray must be float64 with the grid's shape %r" % (plan.shape,))
            cur = Tarray.to(plan._dev).contiguous().clone()
        nxt = torch.empty_like(cur)
        n_rec = nsteps // every
        cells = probe_rec = surf_rec = None
        if probes:
            nz, ny, nx = plan.shape
            for (k, j, i) in probes:
                if not (0 <= k < nz and 0 <= j < ny and 0 <= i < nx):
                    raise IndexError("probe %r outside the grid %r" % ((k, j, i), plan.shape))
            cells = torch.tensor([(k * ny + j) * nx + i for (k, j, i) in probes], dtype=torch.int64, device=plan._dev)
            probe_rec = torch.zeros((n_rec, len(probes)), dtype=torch.float64, device=plan._dev)
        if surface_dz is not None:
            surf_rec = torch.zeros((n_rec,) + plan.shape[1:], dtype=torch.float64, device=plan._dev)
        # bufs[n & 1] is the field before step n
        bufs = [cur, nxt]
        n = 0
        while n < nsteps:
            if plan.source_active(t0 + n * dt, volumetric):
                run_adi_steps(ADI_params, ADI_steps, t0 + n * dt, dt, bufs[n & 1], volumetric_elements, volumetric,
                              out=bufs[(n + 1) & 1])
                n += 1
                if n % every == 0:
                    plan.observe(bufs[n & 1], cells, None if probe_rec is None else probe_rec[n // every - 1],
                                 None if surf_rec is None else surf_rec[n // every - 1], surface_dz)
                continue
            m = n + 1
            while m < nsteps and not plan.source_active(t0 + m * dt, volumetric):
                m += 1
            plan.run_steps_device(bufs[0], bufs[1], n, m - n, every, cells, probe_rec, surf_rec, surface_dz,
                                  first_row=n // every, use_graph=use_graph)
            n = m
        final = bufs[nsteps & 1]
        record = {"step": [(r + 1) * every for r in range(n_rec)]}
        if probes:
            record["probes"] = probe_rec.cpu().numpy()
        if surface_dz is not None:
            record["surface"] = surf_rec.cpu().numpy()
        if was_numpy:
            return plan.download(final), record
        return final, record


def _run_adi_steps_n_stepwise(ADI_params, ADI_steps, t0, dt, Tarray, volumetric_elements, volumetric, nsteps,
                              probes, surface_dz, every):
    """run_adi_steps_n for plans without the native loop (multi-GPU slab plans): one library call per step,
    observation with tensor indexing."""
    import torch
    from . import surface_temperature as st
    plan = ADI_params.plan
    was_numpy = not isinstance(Tarray, torch.Tensor)
    if was_numpy:
        plan.ensure_device()
        cur = torch.from_numpy(np.ascontiguousarray(Tarray, dtype=np.float64)).to(plan._dev)
    else:
        cur = Tarray.contiguous().clone()
    nxt = torch.empty_like(cur)
    rec_steps, rec_probe, rec_surf = [], [], []
    idx = None
    if probes:
        pk, pj, pi = (torch.tensor([p[a] for p in probes], device=cur.device, dtype=torch.long) for a in range(3))
        idx = (pk, pj, pi)
    for n in range(int(nsteps)):
        run_adi_steps(ADI_params, ADI_steps, t0 + n * dt, dt, cur, volumetric_elements, volumetric, out=nxt)
        cur, nxt = nxt, cur
        if (n + 1) % every == 0:
            rec_steps.append(n + 1)
            if idx is not None:
                rec_probe.append(cur[idx])
            if surface_dz is not None:
                rec_surf.append(st.insulating_z_min_surface_temperature(cur, surface_dz))
    if hasattr(plan, "check"):
        plan.check()           # multi-GPU plans: raise if a peer wait ever expired
    record = {"step": rec_steps}
    if idx is not None:
        record["probes"] = torch.stack(rec_probe).cpu().numpy() if rec_probe else np.zeros((0, len(probes)))
    if surface_dz is not None:
        record["surface"] = torch.stack(rec_surf).cpu().numpy() if rec_surf else np.zeros((0,) + tuple(cur.shape[1:]))
    return (cur.cpu().numpy() if was_numpy else cur), record

"""Problem set-up: discretise, classify, and compile to a device plan.

Drop-in for the reference's ``heatsim2/crank_nicolson.pyx``: same ``setup``
signature (:128-142) and the same symbolic helpers ``shift_expression`` (:91)
and ``subst_thermal_conductivity`` (:52).

The reference walks every cell in a Python triple loop (:272-276, 17-26 us and
1.5 kB per cell) although it already de-duplicates the symbolic work through
``expression_cache`` keyed by (boundary ids of the 6 faces, conductivities of
the cell and its 6 neighbours, rho, c) (:340-351).  Here that key is computed
for all cells at once with tensor ops, the distinct keys become *equation
classes*, and the symbolic machinery (boundary plug-ins -> heat balance ->
Crank-Nicolson stage equations) runs once per class.
"""
import numbers
import re

import numpy as np
import torch

from . import alternatingdirection_c_pyx as alternatingdirection
from . import expression
from .plan import AdiPlan

_T_RE = re.compile(r"^T([4-6mp])([4-6mp])([4-6mp])$")
_K_RE = re.compile(r"^kmat([4-6mp])([4-6mp])([4-6mp])$")
_TO_SHIFT = {"4": -1.0, "m": -0.5, "5": 0.0, "p": 0.5, "6": 1.0}
_TO_CHAR = {v: k for k, v in _TO_SHIFT.items()}


def shift_expression(expr, shifts):
    """Re-centre an expression written at a face onto the cell it feeds: every
    ``T***`` / ``kmat***`` variable has its (z,y,x) position digits moved by
    ``shifts`` (multiples of 1/2; 4,m,5,p,6 = -1,-1/2,0,+1/2,+1)."""
    if isinstance(expr, numbers.Number):
        return expr

    def move(name, coef, i0, i1):
        for rx, prefix in ((_T_RE, "T"), (_K_RE, "kmat")):
            m = rx.match(name)
            if m is not None:
                if prefix == "T":
                    assert i0 is None and i1 is None
                digits = "".join(_TO_CHAR[_TO_SHIFT[ch] + s] for ch, s in zip(m.groups(), shifts))
                return ("v", prefix + digits, coef, i0, i1)
        return ("v", name, coef, i0, i1)

    return expr.map_vars(move)


def subst_thermal_conductivity(heatflow_expression, k_params):
    """Insert the conductivities of the cell and its six neighbours, order
    ``(self, x+, y+, z+, x-, y-, z-)`` as in the reference (:52-70, :340)."""
    for name, value in zip(("kmat555", "kmat556", "kmat565", "kmat655", "kmat554", "kmat545", "kmat455"), k_params):
        heatflow_expression = expression.subst(heatflow_expression, name, value)
    return heatflow_expression


def _operand(name):
    return expression.linear_expression(name)


def _face_operands(pos, make=None):
    """Symbolic operands of a face normal to axis ``pos`` (0=z,1=y,2=x) in the
    plug-in argument order: (k-, k+), then (T-, T+) pairs at the face centre
    and displaced by -1/+1 along the first, the second, and both tangential
    axes (reference crank_nicolson.pyx:155-207)."""
    def name(prefix, side, a, b):
        chars = [str(a), str(b)]
        chars.insert(pos, side)
        return prefix + "".join(chars)
    make = _operand if make is None else make
    k = [make(name("kmat", s, 5, 5)) for s in "mp"]
    T = [make(name("T", s, a, b))
         for (a, b) in ((5, 5), (4, 5), (6, 5), (5, 4), (5, 6), (4, 6), (6, 4)) for s in "mp"]
    return k, T


def _plugin_engine(mod):
    """The symbolic engine a boundary plug-in was written against: the module its ``group`` comes from (the
    reference's plug-ins do ``from heatsim2.expression import group``, boundary_conducting.py:5).  None = ours."""
    import sys
    grp = getattr(mod, "group", None)
    eng = sys.modules.get(getattr(grp, "__module__", None)) if grp is not None else None
    if eng is None or eng is expression or not hasattr(eng, "linear_expression"):
        return None
    return eng


def from_foreign_expression(expr):
    """Convert an expression object of the REFERENCE's engine (heatsim2/expression.py: an RPN list ``le_cmdlist`` of
    (operator, value) with OP_ADD/MUL/PARAM/VAR/GROUP/DIV, :15-27) into this package's expression tree, by running the
    RPN program on our own operands.  Lets unmodified reference plug-ins (and user plug-ins written against
    ``heatsim2.expression``) be handed to :func:`setup`."""
    if isinstance(expr, numbers.Number):
        return expr
    if isinstance(expr, expression.linear_expression):
        return expr
    stack = []
    for op, val in expr.le_cmdlist:
        if op == expr.OP_PARAM:
            stack.append(val)
        elif op == expr.OP_VAR:
            name, coef, i0, i1 = val
            v = expression.linear_expression(name)
            if i0 is not None or i1 is not None:
                v = v[i0, i1]
            stack.append(v * coef)
        elif op == expr.OP_GROUP:
            top = stack.pop()
            stack.append(expression.group(top if isinstance(top, expression.linear_expression)
                                          else expression.linear_expression(top)))
        elif op in (expr.OP_ADD, expr.OP_MUL, expr.OP_DIV):
            b = stack.pop()
            a = stack.pop()
            stack.append(a + b if op == expr.OP_ADD else (a * b if op == expr.OP_MUL else a / b))
        else:
            raise ValueError("unknown operator %r in a foreign expression" % (op,))
    if len(stack) != 1:
        raise ValueError("malformed foreign expression")
    res = stack[0]
    return res if isinstance(res, expression.linear_expression) else expression.linear_expression(res)


def evaluate_boundaries(boundaries, dz, dy, dx):
    """Flux expression through a z-, y- and x-face for every boundary class
    (reference :220-250): the plug-in's ``qz/qy/qx`` are called once with
    symbolic operands centred on the face; extra elements of the boundary
    tuple are passed through as trailing arguments.  A plug-in written against
    another engine with the reference's data model (``heatsim2.expression``) gets
    operands of ITS engine and its result is converted (:func:`from_foreign_expression`)."""
    out = []
    for boundary in boundaries:
        mod, extra = boundary[0], list(boundary[1:])
        eng = _plugin_engine(mod)
        fluxes = []
        for pos, fn in enumerate((mod.qz, mod.qy, mod.qx)):
            k, T = _face_operands(pos, None if eng is None else eng.linear_expression)
            fluxes.append(from_foreign_expression(fn(k[0], k[1], dz, dy, dx, *(T + extra))))
        out.append(tuple(fluxes))
    return out


def _as_u8(name, arr, shape):
    a = arr.detach().cpu().numpy() if isinstance(arr, torch.Tensor) else np.asarray(arr)
    if a.dtype != np.uint8:
        raise ValueError("Buffer dtype mismatch, expected 'uint8_t' for %s but got %r" % (name, a.dtype))
    if a.shape != tuple(shape):
        raise ValueError("%s has shape %r, expected %r" % (name, a.shape, tuple(shape)))
    return a


def _cell_keys(material_elements, bz, by, bx, fixed_lut, n_mat, n_bnd, device, extra=None, extra_radix=1):
    """int64 equation key of every cell, computed in z-slabs to bound memory.

    Key digits (mixed radix), mirroring the reference's ``key_params`` (:296,
    :340-347): own material; effective material of the x+,y+,z+,x-,y-,z-
    neighbour (own material when the neighbour is outside the grid or
    TEMPERATURE_FIXED, :298-338); boundary class of the z-,z+,y-,y+,x-,x+
    face.  ``extra`` (int64 tensor [nz,ny,nx] of values < ``extra_radix``) adds a
    last digit - the geometry class of curved-surface mode.  TEMPERATURE_FIXED
    cells all get key -1."""
    nz, ny, nx = material_elements.shape
    radices = [n_mat] * 7 + [n_bnd] * 6 + [int(extra_radix)]
    total = 1
    for r in radices:
        total *= r
    if total >= (1 << 62):
        raise NotImplementedError("too many materials/boundaries for a 62-bit class key")
    if int(material_elements.max()) >= n_mat:
        raise IndexError("material_elements refers to material %d but only %d are defined"
                         % (int(material_elements.max()), n_mat))
    for nm, arr in (("z", bz), ("y", by), ("x", bx)):
        if int(arr.max()) >= n_bnd:
            raise IndexError("boundary_%s_elements refers to boundary %d but only %d are defined"
                             % (nm, int(arr.max()), n_bnd))
    fixed = torch.from_numpy(fixed_lut).to(device)
    keys = torch.empty((nz, ny, nx), dtype=torch.int64, device=device)
    chunk = max(1, int(16e6 // max(1, ny * nx)))
    for k0 in range(0, nz, chunk):
        k1 = min(nz, k0 + chunk)
        # slab with one plane of padding each side; outside the grid the
        # padding repeats the edge plane, i.e. "neighbour = own material"
        planes = [material_elements[max(k0 - 1, 0):max(k0 - 1, 0) + 1] if k0 > 0 else material_elements[0:1],
                  material_elements[k0:k1],
                  material_elements[k1:k1 + 1] if k1 < nz else material_elements[nz - 1:nz]]
        mp = torch.from_numpy(np.concatenate(planes, axis=0)).to(device).long()
        m = mp[1:-1]

        def nbr(axis, sgn):
            """effective neighbour material along axis (0=z,1=y,2=x)"""
            if axis == 0:
                cand = mp[2:] if sgn > 0 else mp[:-2]
            else:
                cand = m.clone()
                src = [slice(None)] * 3
                dst = [slice(None)] * 3
                if sgn > 0:
                    dst[axis], src[axis] = slice(None, -1), slice(1, None)
                else:
                    dst[axis], src[axis] = slice(1, None), slice(None, -1)
                cand[tuple(dst)] = m[tuple(src)]
            return torch.where(fixed[cand], m, cand)

        key = m.clone()
        for axis, sgn in ((2, 1), (1, 1), (0, 1), (2, -1), (1, -1), (0, -1)):
            key = key * n_mat + nbr(axis, sgn)
        zf = torch.from_numpy(bz[k0:k1 + 1]).to(device).long()
        yf = torch.from_numpy(by[k0:k1]).to(device).long()
        xf = torch.from_numpy(bx[k0:k1]).to(device).long()
        for f in (zf[:-1], zf[1:], yf[:, :-1], yf[:, 1:], xf[:, :, :-1], xf[:, :, 1:]):
            key = key * n_bnd + f
        key = key * int(extra_radix)
        if extra is not None:
            key = key + extra[k0:k1].to(device)
        keys[k0:k1] = torch.where(fixed[m], torch.full_like(key, -1), key)
    return keys, radices, total


def _classify(material_elements, bz, by, bx, fixed_lut, n_mat, n_bnd, device, extra=None, extra_radix=1):
    """Per-cell equation key -> dense class ids.  Returns (class_id int32
    tensor [nz,ny,nx], keys numpy int64 [n_classes] ascending, decode)."""
    key, radices, total = _cell_keys(material_elements, bz, by, bx, fixed_lut, n_mat, n_bnd, device, extra, extra_radix)
    flat = key.reshape(-1)
    if total <= (1 << 26):
        # small key space: presence table + prefix sum, no sort
        present = torch.zeros(total + 1, dtype=torch.bool, device=device)
        step = 1 << 26
        for s0 in range(0, flat.numel(), step):
            present[flat[s0:s0 + step] + 1] = True
        class_of_key = (torch.cumsum(present.to(torch.int32), 0) - 1).to(torch.int32)
        class_id = torch.empty(flat.numel(), dtype=torch.int32, device=device)
        for s0 in range(0, flat.numel(), step):
            class_id[s0:s0 + step] = class_of_key[flat[s0:s0 + step] + 1]
        keys = (torch.nonzero(present).reshape(-1) - 1).cpu().numpy()
    else:
        ukeys, inv = torch.unique(flat, return_inverse=True)
        class_id = inv.to(torch.int32)
        keys = ukeys.cpu().numpy()
    class_id = class_id.reshape(key.shape)

    def decode(k):
        vals = []
        for r in reversed(radices):
            vals.append(int(k % r))
            k //= r
        vals.reverse()
        return vals[0], vals[1:7], vals[7:13], vals[13]

    return class_id, keys, decode


def compile_problem(z0, y0, x0,
                    dz, dy, dx,
                    nz, ny, nx,
                    dt,
                    materials,
                    boundaries,
                    volumetric,
                    material_elements,
                    boundary_z_elements,
                    boundary_y_elements,
                    boundary_x_elements,
                    volumetric_elements,
                    top_surface_y_curvatures=None,
                    top_surface_x_curvatures=None,
                    unaligned_anisotropic=False,
                    device=None):
    """Discretise and classify: returns ``(class_id, coefs, volume_array, vol)``
    with ``class_id`` an int32 tensor [nz,ny,nx] of equation-class indices and
    ``coefs`` [n_classes, 8] = (M, gx-,gx+,gy-,gy+,gz-,gz+, D) per class.
    Shared by :func:`setup` (one GPU) and ``dist.setup`` (z-slabs)."""
    from . import TEMPERATURE_COMPUTE, TEMPERATURE_FIXED
    nz, ny, nx = int(nz), int(ny), int(nx)
    curved = top_surface_y_curvatures is not None or top_surface_x_curvatures is not None
    material_elements = _as_u8("material_elements", material_elements, (nz, ny, nx))
    bz = _as_u8("boundary_z_elements", boundary_z_elements, (nz + 1, ny, nx))
    by = _as_u8("boundary_y_elements", boundary_y_elements, (nz, ny + 1, nx))
    bx = _as_u8("boundary_x_elements", boundary_x_elements, (nz, ny, nx + 1))
    vol = _as_u8("volumetric_elements", volumetric_elements, (nz, ny, nx))
    for mat in materials:
        if mat[0] not in (TEMPERATURE_COMPUTE, TEMPERATURE_FIXED):
            raise AssertionError("material type must be TEMPERATURE_COMPUTE or TEMPERATURE_FIXED")
    fixed_lut = np.array([mat[0] == TEMPERATURE_FIXED for mat in materials], dtype=bool)

    evalboundaries = evaluate_boundaries(boundaries, dz, dy, dx)
    volume_array = dz * dy * dx

    work_dev = torch.device("cpu")
    if torch.cuda.is_available():
        work_dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    # Curved-surface mode (reference :388-458): the cell geometry depends on the
    # two curvatures at (j, i) and on the depth index k, and so does the equation
    # (the reference's key_params gain (curv_y, curv_x, k)).  Cells sharing a
    # curvature pair and a depth share a class; the class tables cover up to
    # 65536 classes, i.e. curvature that is constant or piecewise constant over
    # (j, i).  Anything finer needs per-cell coefficients (not implemented).
    extra, extra_radix, curv_pairs = None, 1, None
    if curved:
        cy = np.zeros((ny, nx)) if top_surface_y_curvatures is None else np.asarray(top_surface_y_curvatures, dtype=np.float64)
        cx = np.zeros((ny, nx)) if top_surface_x_curvatures is None else np.asarray(top_surface_x_curvatures, dtype=np.float64)
        if cy.shape != (ny, nx) or cx.shape != (ny, nx):
            raise ValueError("top_surface_*_curvatures must have shape (ny, nx) = %r" % ((ny, nx),))
        curv_pairs, curv_id = np.unique(np.stack([cy.reshape(-1), cx.reshape(-1)], axis=1), axis=0, return_inverse=True)
        curv_id = curv_id.reshape(ny, nx)
        # (more than 65536 (pair, layer) combinations - a curvature MAP on a large grid - still work: 4-byte class
        # ids and the whole-line kernels, plan.py wide_ids; the symbolic pipeline then runs once per class, which is
        # what the reference does per cell)
        extra_radix = len(curv_pairs) * nz
        extra = torch.from_numpy(curv_id[None, :, :] * nz + np.arange(nz)[:, None, None]).to(torch.int64)
        kk = np.arange(nz, dtype=np.float64)[:, None, None]
        dy_mean = 0.5 * ((1.0 + kk * cy[None] * dz) * dy + (1.0 + (kk + 1) * cy[None] * dz) * dy)
        dx_mean = 0.5 * ((1.0 + kk * cx[None] * dz) * dx + (1.0 + (kk + 1) * cx[None] * dz) * dx)
        # like the reference, volume_array stays 0 in TEMPERATURE_FIXED cells (:473-486)
        volume_array = np.where(fixed_lut[material_elements], 0.0, np.abs(dx_mean * dy_mean * dz))
    class_id, keys, decode = _classify(material_elements, bz, by, bx, fixed_lut, len(materials), len(boundaries), work_dev,
                                       extra, extra_radix)
    curved_boundaries = {}

    T555p = _operand("T555p")
    T555m = _operand("T555m")
    volumetric_source = _operand("volumetric_source")
    expression_cache = {}
    coefs = np.zeros((len(keys), 8))
    for c, key in enumerate(keys):
        if key < 0:
            value_key = (TEMPERATURE_FIXED,)
        else:
            mat, nbrs, bnd, geom = decode(int(key))
            (_, matl_k, matl_rho, matl_c) = materials[mat]
            k_params = (matl_k,) + tuple(materials[nb][1] for nb in nbrs)
            k_hash = tuple(tuple(np.ravel(kp)) if isinstance(kp, np.ndarray) else kp for kp in k_params)
            value_key = (TEMPERATURE_COMPUTE, tuple(bnd), k_hash, matl_rho, matl_c)
            if curved:
                kurv_y, kurv_x = (float(v) for v in curv_pairs[geom // nz])
                value_key = value_key + (kurv_y, kurv_x, geom % nz)
        row = expression_cache.get(value_key)
        if row is None:
            if key < 0:
                heatflow = expression.linear_expression(0.0)
                time_expression = -(T555p - T555m)
            elif curved:
                # reference :410-452: areas instead of 1/length, balance times the volume
                (be_m55, be_p55, be_5m5, be_5p5, be_55m, be_55p) = bnd
                kz = geom % nz
                dy_top = (1.0 + kz * kurv_y * dz) * dy
                dy_bot = (1.0 + (kz + 1) * kurv_y * dz) * dy
                dy_m = np.mean((dy_top, dy_bot))
                dx_top = (1.0 + kz * kurv_x * dz) * dx
                dx_bot = (1.0 + (kz + 1) * kurv_x * dz) * dx
                dx_m = np.mean((dx_top, dx_bot))
                volume = np.abs(dx_m * dy_m * dz)
                eb = curved_boundaries.get((dy_m, dx_m))
                if eb is None:      # plug-ins evaluated with this layer's mean edge lengths (reference: symbolic dy, dx)
                    eb = curved_boundaries[(dy_m, dx_m)] = evaluate_boundaries(boundaries, dz, dy_m, dx_m)
                heatflow = (
                    (shift_expression(eb[be_m55][0], (-.5, 0, 0)) * (dx_top * dy_top) - shift_expression(eb[be_p55][0], (+.5, 0, 0)) * (dx_bot * dy_bot)) +
                    (shift_expression(eb[be_5m5][1], (0, -.5, 0)) - shift_expression(eb[be_5p5][1], (0, +.5, 0))) * (dx_m * dz) +
                    (shift_expression(eb[be_55m][2], (0, 0, -.5)) - shift_expression(eb[be_55p][2], (0, 0, +.5))) * (dy_m * dz) +
                    volumetric_source * volume)
                heatflow = subst_thermal_conductivity(heatflow, k_params)
                time_expression = -(T555p - T555m) * matl_rho * matl_c * volume * (1.0 / dt)
            else:
                (be_m55, be_p55, be_5m5, be_5p5, be_55m, be_55p) = bnd
                heatflow = (
                    (shift_expression(evalboundaries[be_m55][0], (-.5, 0, 0)) - shift_expression(evalboundaries[be_p55][0], (+.5, 0, 0))) * (1.0 / dz) +
                    (shift_expression(evalboundaries[be_5m5][1], (0, -.5, 0)) - shift_expression(evalboundaries[be_5p5][1], (0, +.5, 0))) * (1.0 / dy) +
                    (shift_expression(evalboundaries[be_55m][2], (0, 0, -.5)) - shift_expression(evalboundaries[be_55p][2], (0, 0, +.5))) * (1.0 / dx) +
                    volumetric_source)
                heatflow = subst_thermal_conductivity(heatflow, k_params)
                time_expression = -(T555p - T555m) * matl_rho * matl_c * (1.0 / dt)
            (spatial, times) = alternatingdirection.adi_expressions(heatflow, time_expression,
                                                                    unaligned_anisotropic=unaligned_anisotropic)
            eqdicts = alternatingdirection.stage_dicts(spatial, times)
            M, g, D = alternatingdirection.class_coefficients(eqdicts)
            row = expression_cache[value_key] = (M,) + tuple(g) + (D,)
        coefs[c] = row

    # distinct keys can yield the same equation (e.g. a different neighbour
    # behind an insulating face): merge them so the tables stay minimal
    ucoefs, merged = np.unique(coefs, axis=0, return_inverse=True)
    if len(ucoefs) < len(coefs):
        lut = torch.from_numpy(merged.reshape(-1).astype(np.int32)).to(class_id.device)
        class_id = lut[class_id.long()]
        coefs = ucoefs
    return class_id, coefs, volume_array, vol


def setup(z0, y0, x0,
          dz, dy, dx,
          nz, ny, nx,
          dt,
          materials,
          boundaries,
          volumetric,
          material_elements,
          boundary_z_elements,
          boundary_y_elements,
          boundary_x_elements,
          volumetric_elements,
          top_surface_y_curvatures=None,
          top_surface_x_curvatures=None,
          unaligned_anisotropic=False,
          device=None):
    """Compile a heat-conduction problem into ``(ADI_params, ADI_steps)``.

    Arguments as in the reference (crank_nicolson.pyx:128-142).  ``device``
    (extension) picks the CUDA device; default is the current one.  Outer faces
    that would conduct out of the grid raise ValueError (the reference calls
    exit(1), alternatingdirection_c.c:160-163)."""
    nz, ny, nx = int(nz), int(ny), int(nx)
    class_id, coefs, volume_array, vol = compile_problem(
        z0, y0, x0, dz, dy, dx, nz, ny, nx, dt, materials, boundaries, volumetric, material_elements,
        boundary_z_elements, boundary_y_elements, boundary_x_elements, volumetric_elements,
        top_surface_y_curvatures, top_surface_x_curvatures, unaligned_anisotropic, device)
    (ADI_params, ADI_steps) = alternatingdirection.adi_setup((nz, ny, nx), volume_array)
    ADI_params.plan = AdiPlan((nz, ny, nx), class_id, coefs, dt, volume_array, volumetric_elements=vol,
                              materials=materials)
    return (ADI_params, ADI_steps)

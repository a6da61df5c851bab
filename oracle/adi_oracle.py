"""CPU restatement (numpy) of heatsim2's ADI Crank-Nicolson time step.

TEST INFRASTRUCTURE ONLY - this is the parity oracle.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import it; nothing under heatsim2_b200/ does, and the product has no CPU
path.  Parity status: PINNED - tests/test_oracle.py checks this restatement
against (a) the golden vectors in tests/golden/ that were produced by the
unmodified reference built from /root/reference (oracle/build_ref.py,
tests/golden/make_golden.py) and (b) the built reference itself whenever
oracle/_ref is present.

It is written independently of the product's host layer: coefficients come
from closed-form expressions per boundary kind instead of the symbolic
expression engine, every cell carries its own coefficients (no equation
classes, no line tables), the three stages are evaluated in the reference's
literal "direct" form (not the delta form the kernels use) and the Thomas
solve divides by the pivots like heatsim2/tridiag.pyx does.

Reference locations restated here (paths in isuthermography/heatsim2):
  face conductances   heatsim2/boundary_conducting.py:18-25 (mean k / d^2),
                      boundary_thininsulatinglayer.py:19-23 (h / d),
                      boundary_insulating.py:11 (0),
                      boundary_conducting_anisotropic.py:18-24 (K[a,a], aligned)
  neighbour k rule    heatsim2/crank_nicolson.pyx:298-338
  heat balance        heatsim2/crank_nicolson.pyx:359-378
  stage equations     heatsim2/alternatingdirection_c_pyx.pyx:482-493
  matrix entry rules  heatsim2/alternatingdirection_c.c:102-199
  stage loop          heatsim2/alternatingdirection_c_pyx.pyx:389-416
  sources             heatsim2/alternatingdirection_c_pyx.pyx:294-386
  tridiagonal LU      heatsim2/tridiag.pyx:9-43
  tridiagonal solve   heatsim2/tridiag.pyx:46-69
"""
import numpy as np

TEMPERATURE_COMPUTE, TEMPERATURE_FIXED = 0, 1
NO_SOURCE, IMPULSE_SOURCE, STEPPED_SOURCE, IMPULSE_POINT_SOURCE_JOULES, SPATIALLY_Z_DECAYING_TEMPORAL_IMPULSE = range(5)

# axis index into (z, y, x) arrays
Z, Y, X = 0, 1, 2


def _boundary_kind(boundary):
    name = boundary[0].__name__.rsplit(".", 1)[-1]
    kinds = {"boundary_insulating": "insulating", "boundary_conducting": "conducting",
             "boundary_conducting_anisotropic": "anisotropic", "boundary_thininsulatinglayer": "thinlayer"}
    if name not in kinds:
        raise ValueError("oracle knows the four shipped boundary kinds only, got %s" % name)
    return kinds[name]


def _k_axis(k, axis):
    """conductivity seen along an axis: scalar k, or K[a,a] of a tensor"""
    if isinstance(k, np.ndarray):
        return float(k[axis, axis])
    return float(k)


class OracleProblem(object):
    """Per-cell coefficients M, g (6 faces), D of one problem."""

    def __init__(self, z0, y0, x0, dz, dy, dx, nz, ny, nx, dt, materials, boundaries, volumetric,
                 material_elements, boundary_z_elements, boundary_y_elements, boundary_x_elements,
                 volumetric_elements, top_surface_y_curvatures=None, top_surface_x_curvatures=None):
        self.shape = (nz, ny, nx)
        self.dt = dt
        self.d = (dz, dy, dx)
        self.volume = dz * dy * dx
        me = np.asarray(material_elements)
        fixed_of_mat = np.array([m[0] == TEMPERATURE_FIXED for m in materials])
        fixed = fixed_of_mat[me]
        self.fixed = fixed
        rhoc = np.array([1.0 if m[0] == TEMPERATURE_FIXED else m[2] * m[3] for m in materials])
        # capacity term  rho*c*(1/dt)   (crank_nicolson.pyx:378)
        self.M = np.where(fixed, 1.0, rhoc[me] * (1.0 / dt))
        self.D = np.where(fixed, 0.0, 1.0)
        # curved-surface mode (crank_nicolson.pyx:388-458): the balance is multiplied
        # through by the cell volume and every face flux by the face's own area
        curved = top_surface_y_curvatures is not None or top_surface_x_curvatures is not None
        area = None
        if curved:
            kk = np.arange(nz, dtype=np.float64)[:, None, None]
            cy = np.asarray(top_surface_y_curvatures, dtype=np.float64)[None, :, :]
            cx = np.asarray(top_surface_x_curvatures, dtype=np.float64)[None, :, :]
            dy_top, dy_bot = (1.0 + kk * cy * dz) * dy, (1.0 + (kk + 1) * cy * dz) * dy
            dx_top, dx_bot = (1.0 + kk * cx * dz) * dx, (1.0 + (kk + 1) * cx * dz) * dx
            dy_mean, dx_mean = 0.5 * (dy_top + dy_bot), 0.5 * (dx_top + dx_bot)
            vol = np.abs(dx_mean * dy_mean * dz)
            self.volume = np.where(fixed, 0.0, vol)          # volume_array stays 0 in FIXED cells
            self.M = np.where(fixed, 1.0, rhoc[me] * vol * (1.0 / dt))
            self.D = np.where(fixed, 0.0, vol)
            area = {(Z, -1): dx_top * dy_top, (Z, +1): dx_bot * dy_bot,
                    (Y, -1): dx_mean * dz, (Y, +1): dx_mean * dz, (X, -1): dy_mean * dz, (X, +1): dy_mean * dz}
            length = {Z: dz * np.ones(self.shape), Y: dy_mean * np.ones(self.shape), X: dx_mean * np.ones(self.shape)}
        faces = {Z: np.asarray(boundary_z_elements), Y: np.asarray(boundary_y_elements), X: np.asarray(boundary_x_elements)}
        kinds = [_boundary_kind(b) for b in boundaries]
        self.g = {}
        for axis in (Z, Y, X):
            d = self.d[axis]
            k_of_mat = np.array([0.0 if m[0] == TEMPERATURE_FIXED else _k_axis(m[1], axis) for m in materials])
            k_self = k_of_mat[me]
            for side in (-1, +1):
                # neighbour's k, or own k when the neighbour is outside / FIXED
                k_nbr = k_self.copy()
                dst = [slice(None)] * 3
                src = [slice(None)] * 3
                if side > 0:
                    dst[axis], src[axis] = slice(None, -1), slice(1, None)
                else:
                    dst[axis], src[axis] = slice(1, None), slice(None, -1)
                nb_fixed = fixed[tuple(src)]
                k_nbr[tuple(dst)] = np.where(nb_fixed, k_self[tuple(dst)], k_self[tuple(src)])
                # boundary class of that face
                fsl = [slice(None)] * 3
                fsl[axis] = slice(1, None) if side > 0 else slice(None, -1)
                bcls = faces[axis][tuple(fsl)]
                g = np.zeros(self.shape)
                for b, kind in enumerate(kinds):
                    sel = bcls == b
                    if not sel.any() or kind == "insulating":
                        continue
                    if curved:
                        # flux per unit area times the face area (no division by the cell size)
                        if kind in ("conducting", "anisotropic"):
                            val = ((k_self + k_nbr) * 0.5) * (1.0 / length[axis]) * area[(axis, side)]
                        else:
                            val = float(boundaries[b][1]) * area[(axis, side)]
                    elif kind in ("conducting", "anisotropic"):
                        val = ((k_self + k_nbr) * 0.5) * (1.0 / d) * (1.0 / d)
                    else:  # thin layer: q = -h dT, divided by the cell size
                        val = np.full(self.shape, float(boundaries[b][1]) * (1.0 / d))
                    g[sel] = val[sel]
                g[fixed] = 0.0
                self.g[(axis, side)] = g
        # closed outer faces (alternatingdirection_c.c:160-163 exits otherwise)
        for axis in (Z, Y, X):
            lo = [slice(None)] * 3
            hi = [slice(None)] * 3
            lo[axis], hi[axis] = 0, -1
            if self.g[(axis, -1)][tuple(lo)].any() or self.g[(axis, +1)][tuple(hi)].any():
                raise ValueError("Equation exceeds bounds of domain. Are external boundaries set correctly?")
        self.volumetric = volumetric
        self.volumetric_elements = np.asarray(volumetric_elements)
        self._lu = {}

    # ---------------------------------------------------------------- sources
    def volumetric_array(self, t, dt, volumetric_elements=None, volumetric=None):
        """alternatingdirection_c_pyx.pyx:294-386"""
        ve = self.volumetric_elements if volumetric_elements is None else np.asarray(volumetric_elements)
        volumetric = self.volumetric if volumetric is None else volumetric
        out = np.zeros(self.shape)
        for idx, entry in enumerate(volumetric):
            kind = entry[0]
            if kind == IMPULSE_SOURCE:
                if t == entry[1]:
                    out[ve == idx] = entry[2] / dt
            elif kind == STEPPED_SOURCE:
                if t >= entry[1] and t <= entry[2]:
                    out[ve == idx] = entry[3]
            elif kind == IMPULSE_POINT_SOURCE_JOULES:
                if t == entry[1]:
                    sel = ve == idx
                    out[sel] = entry[2] / ((self.volume[sel] if np.ndim(self.volume) else self.volume) * dt)
            elif kind == SPATIALLY_Z_DECAYING_TEMPORAL_IMPULSE:
                (_, t_imp, direc, offset, z_ndgrid, ddz, jpm2, clen) = entry
                if t == t_imp:
                    ddz = abs(float(ddz))
                    centre = np.asarray(z_ndgrid) * direc[0] - float(offset)
                    left = centre - ddz / 2.0
                    right = centre + ddz / 2.0
                    use = (ve == idx) & (right > 0.0)
                    r = right[use]
                    l = left[use]
                    l[l < 0.0] = 0.0
                    zint = -np.exp(-r / float(clen)) + np.exp(-l / float(clen))
                    out[use] = zint * float(jpm2) / (ddz * dt)
        return out

    # ------------------------------------------------------------ operators
    def _L(self, axis, T):
        """L_a T = g-(T[-1]-T) + g+(T[+1]-T), evaluated as the reference's
        B/C rows do: sum of coefficient*value products."""
        gm, gp = self.g[(axis, -1)], self.g[(axis, +1)]
        out = -(gm + gp) * T
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[axis], hi[axis] = slice(None, -1), slice(1, None)
        out[tuple(hi)] += gm[tuple(hi)] * T[tuple(lo)]
        out[tuple(lo)] += gp[tuple(lo)] * T[tuple(hi)]
        return out

    def _solve(self, axis, rhs):
        """(M - L_a/2) x = rhs along every line of ``axis``: LU as
        tridiag.pyx:9-43, forward/back substitution as :46-69."""
        if axis not in self._lu:
            gm, gp = self.g[(axis, -1)], self.g[(axis, +1)]
            a = np.moveaxis(-0.5 * gm, axis, 0)
            b = np.moveaxis(self.M + 0.5 * (gm + gp), axis, 0)
            c = np.moveaxis(-0.5 * gp, axis, 0)
            n = a.shape[0]
            piv = np.empty_like(b)
            u2 = np.empty_like(b)
            piv[0] = b[0]
            for r in range(n):
                u2[r] = c[r] / piv[r]
                if r < n - 1:
                    piv[r + 1] = b[r + 1] - u2[r] * a[r + 1]
            self._lu[axis] = (a.copy(), piv, u2)
        a, piv, u2 = self._lu[axis]
        bvec = np.moveaxis(rhs, axis, 0)
        n = bvec.shape[0]
        x = np.empty_like(bvec)
        x[0] = bvec[0] / piv[0]
        for r in range(1, n):
            x[r] = (bvec[r] - a[r] * x[r - 1]) / piv[r]
        for r in range(n - 2, -1, -1):
            x[r] = x[r] - u2[r] * x[r + 1]
        return np.ascontiguousarray(np.moveaxis(x, 0, axis))

    # ----------------------------------------------------------------- step
    def step(self, t, dt, T, volumetric_elements=None, volumetric=None):
        """One time step, the reference's three stage equations in direct form
        (alternatingdirection_c_pyx.pyx:389-416, 482-493)."""
        T = np.asarray(T, dtype=np.float64)
        s = self.D * self.volumetric_array(t, dt, volumetric_elements, volumetric)
        Lx, Ly, Lz = self._L(X, T), self._L(Y, T), self._L(Z, T)
        MT = self.M * T
        # stage 0: (M - Lx/2) T* = (M + Lx/2 + Ly + Lz) T + s
        T1 = self._solve(X, MT + 0.5 * Lx + Ly + Lz + s)
        # stage 1: (M - Ly/2) T** = (M + Lx/2 + Ly/2 + Lz) T + Lx/2 T* + s
        hLx1 = 0.5 * self._L(X, T1)
        T2 = self._solve(Y, MT + 0.5 * Lx + 0.5 * Ly + Lz + hLx1 + s)
        # stage 2: (M - Lz/2) T' = (M + (Lx+Ly+Lz)/2) T + Lx/2 T* + Ly/2 T** + s
        hLy2 = 0.5 * self._L(Y, T2)
        return self._solve(Z, MT + 0.5 * (Lx + Ly + Lz) + hLx1 + hLy2 + s)


def setup(*args):
    """Same positional arguments as heatsim2.setup (crank_nicolson.pyx:128-142),
    including the two optional top-surface curvature arrays."""
    return OracleProblem(*args)


def run(problem_dict, nsteps=None, record=None):
    """Run a tests/problems.py problem on the oracle.  Returns the final field
    (and, with ``record`` = iterable of step counts, a dict step -> field)."""
    P = setup(*problem_dict["setup_args"])
    T = np.array(problem_dict["T0"], dtype=np.float64)
    dt, t0 = problem_dict["dt"], problem_dict["t0"]
    n = problem_dict["nsteps"] if nsteps is None else nsteps
    rec = {}
    for it in range(n):
        T = P.step(t0 + dt * it, dt, T)
        if record is not None and (it + 1) in record:
            rec[it + 1] = T.copy()
    return (T, rec) if record is not None else T


# --------------------------------------------------------------------- tridiag
def tridiaglu(Amat):
    """heatsim2/tridiag.pyx:9-43"""
    A = np.asarray(Amat, dtype=np.float64)
    n = A.shape[0]
    L = np.zeros((n, 3))
    U = A.copy()
    assert U[0, 0] == 0.0 and U[-1, 2] == 0.0
    for row in range(n):
        L[row, 1] = U[row, 1]
        U[row, 2] /= U[row, 1]
        U[row, 1] = 1.0
        if row < n - 1:
            coefficient = U[row + 1, 0]
            U[row + 1, 0] = 0.0
            U[row + 1, 1] -= U[row, 2] * coefficient
            L[row + 1, 0] = coefficient
    return L, U


def tridiagsolve(Lmat, Umat, bvec):
    """heatsim2/tridiag.pyx:46-69"""
    n = len(bvec)
    x = np.zeros(n)
    x[0] = bvec[0] / Lmat[0, 1]
    for row in range(1, n):
        x[row] = (bvec[row] - Lmat[row, 0] * x[row - 1]) / Lmat[row, 1]
    for row in range(n - 2, -1, -1):
        x[row] = x[row] - Umat[row, 2] * x[row + 1]
    return x

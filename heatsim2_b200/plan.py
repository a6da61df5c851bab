"""AdiPlan - everything the CUDA kernels need for one problem, and the host
logic that drives them.

A plan is built by ``crank_nicolson.setup`` from
  * ``class_id``   per-cell equation-class index (u8/u16 array [nz,ny,nx]),
  * ``class_coef`` per-class ``(M, gx-,gx+,gy-,gy+,gz-,gz+, D)``,
and derives, per sweep axis, the set of *unique tridiagonal lines* and their
Thomas factors.  The reference keeps ``Amat/Lmat/Umat`` (9 doubles per cell per
stage, heatsim2/tridiag.pyx:9-43); all BASELINE grids have a handful of unique
lines per axis, so the factor tables here are a few kB and live in L1/L2.

torch tensors are used as device buffers only; every kernel launch goes through
the C ABI (``_cabi``).  Nothing here computes a time step on the CPU.
"""
import ctypes
import os

import numpy as np
import torch

from . import _cabi

# column order of class_coef rows kept on the host (un-scaled, as parsed)
M_, GXM, GXP, GYM, GYP, GZM, GZP, D_ = range(8)

_AXIS_G = ((GXM, GXP), (GYM, GYP), (GZM, GZP))     # axis 0=x, 1=y, 2=z


def _line_view(a3, axis):
    """View a [nz,ny,nx] tensor as [n_lines, L] rows for sweep axis 0=x,1=y,2=z
    with the library's line numbering (x: k*ny+j, y: k*nx+i, z: j*nx+i)."""
    nz, ny, nx = a3.shape
    if axis == 0:
        return a3.reshape(nz * ny, nx)
    if axis == 1:
        return a3.permute(0, 2, 1).reshape(nz * nx, ny)
    return a3.permute(1, 2, 0).reshape(ny * nx, nz)


def slab_chunk(nz_local):
    """Chunk size of the distributed z-solve: the largest of 32/16/8 that
    divides the slab thickness with at most 32 chunks per slab."""
    forced = int(os.environ.get("HS2_SLAB_CHUNK", "0"))
    for M in ((forced,) if forced in (8, 16, 32) else (32, 16, 8)):
        if nz_local % M == 0 and nz_local // M <= 32:
            return M
    raise NotImplementedError("slab thickness %d must be a multiple of 8 (at most 32 chunks of 8/16/32 planes)" % nz_local)


def thomas_factors(lo, dg, hi):
    """Thomas factorisation of a batch of tridiagonal lines [n_unique, L]:
    returns [n_unique, L, 4] = {1/pivot, lo/pivot, hi/pivot, 0}
    (same recurrence as heatsim2/tridiag.pyx:25-41, stored as reciprocals)."""
    nu, L = dg.shape
    out = np.zeros((nu, L, _cabi.HS2_LU_STRIDE))
    cp_prev = np.zeros(nu)
    for r in range(L):
        piv = dg[:, r] - lo[:, r] * cp_prev
        inv = 1.0 / piv
        cp_prev = hi[:, r] * inv
        out[:, r, 0] = inv
        out[:, r, 1] = lo[:, r] * inv
        out[:, r, 2] = cp_prev
    return out


XW_HDR = 128          # doubles per unique line in front of its interface rows (hs2_axis_tables.d_xw_tab)


def x_warp_applies(nx, n_classes=None):
    """The warp-per-line x sweep (csrc/kernels_xw.cu) takes lines of 16, 32 or 64 chunks of 16 cells (two lines
    per warp, one, or two warps per line) and up to 64 equation classes; HS2_X_KERNEL=tma keeps the patch
    kernel, =fold the LSU-fed one.  Lines of 1024 cells have no other kernel for 64 chunks of 16, so there the
    class count decides the chunking too."""
    if os.environ.get("HS2_X_KERNEL", "warp") != "warp":
        return False
    return nx in (256, 512) or (nx == 1024 and (n_classes is None or n_classes <= 64))


def ghost_uniform_tables(lo, dg, hi, M, tol=1e-16):
    """Tables of the warp-per-line x kernel for the unique lines that have constant coefficients and closed ends:
    rows 1..L-2 identical (a, b, c), row 0 = (0, b + a, c), row L-1 = (a, b + c, 0) (to 4 ulp) - a homogeneous line
    between two insulated / closed faces.  Such a line equals the infinite constant-coefficient line with MIRRORED
    ghost neighbours (x_{-1} = x_0, x_L = x_{L-1}), so that every one of its chunks, the first and the last included,
    can be eliminated with ONE chunk table; the mirror only enters the interface system (alpha_0 = F_0, the first
    value of the line; F_P = E_{P-1}) and the back substitution of the first chunk, x = a - F_0 b with
    b_k = s_k - cp_k b_{k+1} and F_0 = a_0 / (1 + b_0).

    lo, dg, hi: [n_unique, L].  Returns (code uint8 [n_unique], xw [n_unique, XW_HDR + (2 band + 1) * 64], band):
    xw[u, 0:80] planes inv, f, c, s, cp of the common chunk, xw[u, 80:96] b, xw[u, 96] 1 / (1 + b_0), then
    [2 band + 1][P][2] the rows of the inverse interface operator that give E_p, relative to the diagonal
    (entry d of chunk p multiplies (y_first, y_last) of chunk p + d - band; 0 outside the line); band: half-width
    beyond which all entries are below ``tol`` of their row maximum."""
    nu, L = dg.shape
    P = L // M
    code = np.zeros(nu, dtype=np.uint8)
    if L != P * M or L < 3 * M:
        return code, np.zeros((nu, XW_HDR + 64)), 0
    a, b, c = lo[:, 1], dg[:, 1], hi[:, 1]
    eps = 4 * np.finfo(float).eps
    ok = (lo[:, 1:] == a[:, None]).all(axis=1) & (hi[:, :-1] == c[:, None]).all(axis=1) & \
         (dg[:, 1:-1] == b[:, None]).all(axis=1) & (lo[:, 0] == 0.0) & (hi[:, -1] == 0.0) & \
         (np.abs(dg[:, 0] - (b + a)) <= eps * np.abs(b)) & (np.abs(dg[:, -1] - (b + c)) <= eps * np.abs(b))
    code[:] = ok
    sel = np.nonzero(ok)[0]
    if len(sel) == 0:
        return code, np.zeros((nu, XW_HDR + 64)), 0
    ns = len(sel)
    # the common chunk: rows (a, b, c) throughout, coupled to its left neighbour through row 0
    tab, _ = chunk_factors(np.repeat(a[sel, None], M, 1), np.repeat(b[sel, None], M, 1), np.repeat(c[sel, None], M, 1), M)
    inv, f, cc, s, cp = (tab[:, pl, :M] for pl in range(T_PLANES))
    f = f.copy()
    f[:, 0] = 0.0
    v0 = np.zeros(ns)
    c_cur = np.ones(ns)
    for r in range(M):
        v0 += c_cur * s[:, r]
        c_cur = -cp[:, r] * c_cur
    cm, sl, cpl = c_cur, s[:, M - 1], cp[:, M - 1]
    R = np.zeros((ns, 2 * P, 2 * P))
    idx = np.arange(P)
    R[:, idx, idx] = 1.0
    R[:, P + idx, P + idx] = 1.0
    for p in range(P):
        if p > 0:
            R[:, p, P + p - 1] = v0
            R[:, P + p, P + p - 1] = sl
        else:                       # mirrored ghost: alpha_0 = F_0
            R[:, 0, 0] += v0
            R[:, P, 0] += sl
        if p < P - 1:
            R[:, p, p + 1] = -cm
            R[:, P + p, p + 1] = cpl
        else:                       # mirrored ghost: F_P = E_{P-1}
            R[:, p, P + p] += -cm
            R[:, P + p, P + p] += cpl
    Rinv = np.linalg.inv(R)
    GE = np.zeros((ns, P, 2 * P))
    GE[:, :, 0::2] = Rinv[:, P:, :P]
    GE[:, :, 1::2] = Rinv[:, P:, P:]
    band = interface_band(GE, tol)
    bt = np.zeros((ns, M))
    for k in range(M - 2, -1, -1):
        bt[:, k] = s[:, k] - cp[:, k] * bt[:, k + 1]
    w = 2 * band + 1
    xw = np.zeros((nu, XW_HDR + w * P * 2))
    hdr = np.zeros((ns, XW_HDR))
    for pl, t in enumerate((inv, f, cc, s, cp, bt)):
        hdr[:, pl * M:(pl + 1) * M] = t
    hdr[:, 6 * M] = 1.0 / (1.0 + bt[:, 0])
    rel = np.zeros((ns, w, P, 2))
    for d in range(w):
        q = idx + d - band
        inside = (q >= 0) & (q < P)
        rel[:, d, idx[inside], 0] = GE[:, idx[inside], 2 * q[inside]]
        rel[:, d, idx[inside], 1] = GE[:, idx[inside], 2 * q[inside] + 1]
    xw[sel, :XW_HDR] = hdr
    xw[sel, XW_HDR:] = rel.reshape(ns, -1)
    return code, xw, band


def x_tma_applies(nx):
    """The TMA-fed x sweep (csrc/kernels_xt.cu) takes lines of 16..512 cells, a multiple
    of 16 (one 128-byte segment per thread); HS2_X_KERNEL=fold keeps the LSU-fed kernel."""
    return os.environ.get("HS2_X_KERNEL", "warp") != "fold" and nx % 16 == 0 and 16 <= nx <= 512


def choose_chunk(L, axis=None, n_classes=None):
    """Rows per chunk M and chunk count P for a line of length L.  The kernels
    hold M doubles per thread in registers and run P*W threads per tile
    (W = 16 or 8 adjacent lines): M=8 with up to 16 chunks, M=16 with up to 32,
    M=32 with up to 32.  Measured on B200, longer chunks win as long as a line
    still has >= 4 of them, so the largest valid M <= HS2_CHUNK (default 32;
    HS2_CHUNK_X for the x axis) with at least 4 chunks is taken, else the
    largest valid M.  (0, 0): too long for the register-tile kernels
    (whole-line fallback)."""
    if axis == 0 and (x_tma_applies(L) or x_warp_applies(L, n_classes)):
        return 16, L // 16                 # kernels_xt.cu / kernels_xw.cu: one 128-byte segment per thread
    valid = [(M, -(-L // M)) for M, cap in ((8, 16), (16, 32), (32, 32)) if -(-L // M) <= cap]
    if not valid:
        return 0, 0
    pref = int(os.environ.get("HS2_CHUNK", "32"))
    if axis == 0:
        pref = int(os.environ.get("HS2_CHUNK_X", str(pref)))
    good = [(M, P) for M, P in valid if M <= pref and P >= 4]
    if good:
        return good[-1]
    small = [(M, P) for M, P in valid if M <= pref]
    return small[0] if small else valid[0]


T_INV, T_F, T_C, T_S, T_CP, T_PLANES = 0, 1, 2, 3, 4, 5


def chunk_factors(lo, dg, hi, M):
    """Tables of the partitioned (SPIKE-type) tridiagonal solve for a batch of
    unique lines [n_unique, L] cut into chunks of M rows.

    Per chunk, with the coupling to the neighbouring chunks moved to the right
    hand side (alpha = x just before the chunk, beta = x just after it):
      forward   u_k = d_k*inv_k - f_k*u_{k-1}            (u_{-1} = 0)
      first row y_f = sum_k c_k u_k ,  last row y_l = u_last
      backward  x_k = (u_k - alpha*s_k) - cp_k*x_{k+1}   (x_last = E known)
    tab[nu, 5, pitch]: planes inv, f, c, s, cp (pitch = L rounded up to 4).
    The first/last values (F_p, E_p) of all chunks of a line solve a 2P x 2P
    system whose matrix depends on the line only; GE[nu, P, 2P] holds, for
    every chunk, the row of its inverse that gives E_p from the interleaved
    right hand side (y_f0, y_l0, y_f1, y_l1, ...).  alpha_p = E_{p-1}."""
    nu, L = dg.shape
    P = -(-L // M)
    pitch = -(-L // 4) * 4
    tab = np.zeros((nu, T_PLANES, pitch))
    tab[:, T_INV, :] = 1.0
    v0 = np.zeros((nu, P))
    cm = np.zeros((nu, P))
    sl = np.zeros((nu, P))
    cpl = np.zeros((nu, P))
    for p in range(P):
        r0, r1 = p * M, min(L, (p + 1) * M)
        cp_prev = np.zeros(nu)
        s_prev = np.zeros(nu)
        c_cur = np.ones(nu)
        acc = np.zeros(nu)
        for r in range(r0, r1):
            piv = dg[:, r] - (lo[:, r] * cp_prev if r > r0 else 0.0)
            inv = 1.0 / piv
            f = lo[:, r] * inv
            cp = hi[:, r] * inv
            s = f if r == r0 else -f * s_prev
            tab[:, T_INV, r] = inv
            tab[:, T_F, r] = f if r > r0 else 0.0    # the kernel's recurrence restarts at a chunk start
            tab[:, T_C, r] = c_cur
            tab[:, T_S, r] = s
            tab[:, T_CP, r] = cp
            acc += c_cur * s
            c_cur = -cp * c_cur
            cp_prev, s_prev = cp, s
        v0[:, p], cm[:, p], sl[:, p], cpl[:, p] = acc, c_cur, s_prev, cp_prev
    # reduced system, unknown order [F_0..F_{P-1}, E_0..E_{P-1}]:
    #   F_p + v0_p E_{p-1} - cm_p F_{p+1} = yf_p ;  E_p + sl_p E_{p-1} + cpl_p F_{p+1} = yl_p
    R = np.zeros((nu, 2 * P, 2 * P))
    idx = np.arange(P)
    R[:, idx, idx] = 1.0
    R[:, P + idx, P + idx] = 1.0
    for p in range(P):
        if p > 0:
            R[:, p, P + p - 1] = v0[:, p]
            R[:, P + p, P + p - 1] = sl[:, p]
        if p < P - 1:
            R[:, p, p + 1] = -cm[:, p]
            R[:, P + p, p + 1] = cpl[:, p]
    Rinv = np.linalg.inv(R)
    GE = np.zeros((nu, P, 2 * P))
    GE[:, :, 0::2] = Rinv[:, P:, :P]
    GE[:, :, 1::2] = Rinv[:, P:, P:]
    return tab, GE


def interleave_chunks(tab, L, M, P):
    """Chunk-interleaved copy of the factor tables for the x-sweep kernel
    (kernels_xf.cu): [nu, planes, M/2, P, 2] with rows 2t, 2t+1 of chunk p at
    [.., t, p, :].  A warp of that kernel works on four neighbouring chunks of
    eight lines; this order puts the four table entries it loads together in one
    64-byte segment.  Rows past the end of the line: 1/piv = 1, everything else 0."""
    nu = tab.shape[0]
    full = np.zeros((nu, T_PLANES, P * M))
    full[:, T_INV, :] = 1.0
    full[:, :, :L] = tab[:, :, :L]
    return np.ascontiguousarray(full.reshape(nu, T_PLANES, P, M // 2, 2).transpose(0, 1, 3, 2, 4))


def uniform_chunks(tab, L, M, P, line_id=None):
    """The most common chunk table of an axis and where it applies.

    The chunk-local factorisation of ``chunk_factors`` restarts at every chunk
    start, so every interior chunk of a line with constant coefficients - and of
    any other line that crosses the same material there - has bit-identical
    tables.  Returns ``(utab [T_PLANES, M], ucode uint8 [n_unique, P])`` with
    ``ucode[u, p] = 1`` where the full chunk p of unique line u equals ``utab``
    bit for bit; chunks are weighted by the number of lines that use them
    (``line_id``: int array/tensor of unique-line ids, optional).  The kernels
    get ``utab`` by value and read it as constant operands (chunk_core.cuh)."""
    nu = tab.shape[0]
    n_full = L // M
    ucode = np.zeros((nu, P), dtype=np.uint8)
    if n_full == 0:
        return np.zeros((T_PLANES, M)), ucode
    if line_id is None:
        weight = np.ones(nu, dtype=np.int64)
    else:
        lid = line_id.cpu().numpy() if isinstance(line_id, torch.Tensor) else np.asarray(line_id)
        weight = np.bincount(lid.astype(np.int64), minlength=nu)
    blocks = np.ascontiguousarray(tab[:, :, :n_full * M].reshape(nu, T_PLANES, n_full, M).transpose(0, 2, 1, 3))
    flat = blocks.reshape(nu * n_full, T_PLANES * M)
    keys = flat.view(np.dtype((np.void, flat.dtype.itemsize * flat.shape[1]))).reshape(-1)
    uniq, inverse = np.unique(keys, return_inverse=True)
    votes = np.bincount(inverse.reshape(-1), weights=np.repeat(weight, n_full).astype(np.float64), minlength=len(uniq))
    best = int(np.argmax(votes))
    first = int(np.nonzero(inverse.reshape(-1) == best)[0][0])
    utab = flat[first].reshape(T_PLANES, M).copy()
    ucode[:, :n_full] = (inverse.reshape(nu, n_full) == best)
    return utab, ucode


def interface_band(GE, tol=1e-16):
    """Half-width, in chunks, outside which every entry of every GE row is
    below ``tol`` times the row maximum (the interface operator decays
    geometrically away from the diagonal)."""
    nu, P, _ = GE.shape
    mag = np.maximum(np.abs(GE[:, :, 0::2]), np.abs(GE[:, :, 1::2]))        # [nu, P(row), P(col)]
    rowmax = mag.max(axis=2, keepdims=True)
    dist = np.abs(np.arange(P)[:, None] - np.arange(P)[None, :])
    sig = mag > tol * rowmax
    return int((dist[None, :, :] * sig).max()) if sig.any() else 0


class AdiPlan(object):
    """One problem (or one z-slab of it) compiled for the CUDA library.

    ``slab = (k0, class_id_global)`` makes this the plan of planes
    ``[k0, k0 + shape[0])`` of a larger grid split along z over several GPUs:
    x- and y-line tables come from the local planes, z-line tables from the
    global lines (the z-solve spans all slabs, see ``dist.py``)."""

    MAX_TABLE_SOURCES = 8      # hs2_source table form (x_common.cuh SrcTab)

    def __init__(self, shape, class_id, class_coef, dt, volume_array, volumetric_elements=None,
                 materials=None, slab=None):
        self.shape = tuple(int(s) for s in shape)
        nz, ny, nx = self.shape
        self.n = nz * ny * nx
        self.dt = dt
        self.volume_array = volume_array
        self.materials = materials
        self.class_coef = np.ascontiguousarray(class_coef, dtype=np.float64)     # [nc, 8] un-scaled
        self.n_classes = self.class_coef.shape[0]
        if self.n_classes >= (1 << 31):
            raise NotImplementedError("more than 2^31 equation classes (%d)" % self.n_classes)
        # more than 65536 classes (per-cell equations: a curvature MAP in curved-surface mode), or HS2_WIDE_IDS=1:
        # 4-byte class ids, whole-line kernels (the class table does not fit shared memory, every line is unique)
        self.wide_ids = self.n_classes > 65536 or os.environ.get("HS2_WIDE_IDS", "0") == "1"
        if self.wide_ids and slab is not None:
            raise NotImplementedError("z-slab plans need at most 65536 equation classes (%d)" % self.n_classes)
        self.slab = None
        if slab is None:
            self.class_id = self._pack_ids(class_id).reshape(self.shape).contiguous()
            self._check_closed(self.class_id)
        else:
            k0, gid = slab
            gid = self._pack_ids(gid)
            self._check_closed(gid)
            if k0 < 0 or k0 + nz > gid.shape[0] or tuple(gid.shape[1:]) != (ny, nx):
                raise ValueError("slab [%d,%d) does not fit the global grid %r" % (k0, k0 + nz, tuple(gid.shape)))
            self.slab = dict(k0=int(k0), nz_global=int(gid.shape[0]))
            self._global_ids = gid
            self.class_id = gid[k0:k0 + nz].contiguous()
        self._setup_vol_elements = volumetric_elements
        if self.slab is None:
            self._global_ids = None
        self._np = None               # host (numpy) statement of the tables, built on first use: tests, emulation
        self._dev = None
        self._handle = None
        # HS2_FORCE_FALLBACK=1: run the whole-line global-memory kernels (testing aid)
        self.flags = 1 if os.environ.get("HS2_FORCE_FALLBACK", "0") == "1" else 0
        # HS2_NO_UTAB=1: every chunk reads its factor tables (A/B of the constant-bank fast path)
        if os.environ.get("HS2_NO_UTAB", "0") == "1":
            self.flags |= 4
        if os.environ.get("HS2_X_KERNEL", "warp") == "fold":
            self.flags |= 16
        if os.environ.get("HS2_X_KERNEL", "warp") == "tma":
            self.flags |= 32
        # axes whose kernels take the constant-bank table (measured on B200, profiles/NOTES_r02.md)
        self.utab_axes = os.environ.get("HS2_UTAB_AXES", "xz")     # x: the TMA-fed kernel; z: strided_sweep_tma<FINAL>
        self._bufs = {}
        self._vol_dev = None
        self._vol_key = None

    def _pack_ids(self, class_id):
        cid = class_id if isinstance(class_id, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(class_id))
        if self.wide_ids:
            return cid.to(torch.int32)
        want = torch.uint8 if self.n_classes <= 256 else torch.int16    # int16 carries u16 bit patterns
        if cid.dtype != want:
            cid = cid.to(torch.int64).to(want) if want == torch.uint8 else cid.to(torch.int32).to(torch.int16)
        return cid

    # ------------------------------------------------------------ validation
    def _check_closed(self, cid):
        """No conductance may point out of the domain.  The reference detects
        this inside C add_equation and calls exit(1)
        (alternatingdirection_c.c:160-163); here it is a ValueError."""
        faces = ((cid[0], GZM, "z-min"), (cid[-1], GZP, "z-max"),
                 (cid[:, 0], GYM, "y-min"), (cid[:, -1], GYP, "y-max"),
                 (cid[:, :, 0], GXM, "x-min"), (cid[:, :, -1], GXP, "x-max"))
        for sl, col, name in faces:
            ids = torch.unique(sl.to(torch.int32) & (0xFFFF if sl.dtype == torch.int16 else 0x7FFFFFFF)).cpu().numpy()
            if np.any(self.class_coef[ids, col] != 0.0):
                raise ValueError("Equation exceeds bounds of domain on the %s face. "
                                 "Are external boundaries set correctly?" % name)

    # ------------------------------------------------------- unique line tables
    # The product path builds every table natively (hs2_plan_build, csrc/plan_build.cu).  What follows is the numpy
    # statement of the same algebra: the tests compare the two, the CPU emulation of the multi-GPU solve and
    # reference_matrices read it, and HS2_TABLES=numpy feeds it to hs2_plan_create for A/B runs.
    @property
    def scaled_coef(self):
        cc = self.class_coef
        cap = cc[:, M_]
        out = np.zeros((self.n_classes, _cabi.HS2_COEF_STRIDE))
        out[:, 0:6] = cc[:, GXM:GZP + 1] / cap[:, None]
        out[:, 6] = cc[:, D_] / cap
        out[:, 7] = cap
        return out

    @property
    def chunk(self):
        """(rows per chunk, chunks) of the partitioned solve per axis - policy only, no tables"""
        if self._handle is not None:
            return [(i.chunk, i.n_chunks) for i in (self.axis_info(a) for a in range(3))]
        out = []
        if self.wide_ids:
            return [(0, 0)] * 3
        for axis in range(3):
            if axis == 2 and self.slab is not None:
                M = slab_chunk(self.shape[0])
                out.append((M, self.slab["nz_global"] // M))
            else:
                out.append(choose_chunk(self.shape[2 - axis], axis, self.n_classes))
        return out

    def axis_info(self, axis):
        """hs2_plan_axis_info of the device plan"""
        info = _cabi.AxisInfo()
        _cabi.check(_cabi.lib().hs2_plan_axis_info(self._handle, axis, ctypes.byref(info)))
        return info

    def copy_table(self, axis, which, dtype=np.float64):
        """host copy of a table of the natively built plan (hs2_plan_copy_table), flat"""
        lib = _cabi.lib()
        n = int(lib.hs2_plan_copy_table(self._handle, axis, which, None, 0))
        if n < 0:
            _cabi.check(n)
        out = np.empty(n // np.dtype(dtype).itemsize, dtype=dtype)
        if n:
            got = int(lib.hs2_plan_copy_table(self._handle, axis, which, out.ctypes.data, n))
            if got < 0:
                _cabi.check(got)
        return out

    def _tables_np(self):
        if self._np is None:
            self._np = self._build_lines()
        return self._np

    line_id = property(lambda self: self._tables_np()["line_id"])
    line_lu = property(lambda self: self._tables_np()["line_lu"])
    line_rows = property(lambda self: self._tables_np()["line_rows"])
    chunk_tabs = property(lambda self: self._tables_np()["chunk_tabs"])

    def _build_lines(self):
        cc = self.class_coef
        cap = cc[:, M_]
        if self.slab is not None and self._global_ids is None:
            raise RuntimeError("the global class ids of this slab plan were released after the native build")

        def as_int(cid):
            return (cid.to(torch.int32) & 0xFFFF) if cid.dtype == torch.int16 else cid

        line_id, line_lu, line_rows, chunk_tabs = [], [], [], []
        policy = self.chunk if self._handle is None else None
        for axis in range(3):
            gm, gp = _AXIS_G[axis]
            rows = np.stack([-0.5 * cc[:, gm] / cap, 1.0 + 0.5 * (cc[:, gm] + cc[:, gp]) / cap, -0.5 * cc[:, gp] / cap], axis=1)
            urows, sub_of_class = np.unique(rows, axis=0, return_inverse=True)
            sub_of_class = sub_of_class.reshape(-1)
            cid = self._global_ids if (axis == 2 and self.slab is not None) else self.class_id
            sub_lut = torch.from_numpy(sub_of_class.astype(np.int64)).to(cid.device)
            lid, reps = _unique_lines(as_int(cid), sub_lut, axis, len(urows))
            lo, dg, hi = (urows[:, c][reps] for c in range(3))
            line_id.append(lid)                       # int32 tensor [n_lines]
            line_rows.append((lo, dg, hi))            # numpy [n_unique, L] each
            if axis == 2 and self.slab is not None:
                line_lu.append(np.zeros((dg.shape[0], 1, _cabi.HS2_LU_STRIDE)))    # whole-line path unused
            else:
                line_lu.append(thomas_factors(lo, dg, hi))
            rows_per_chunk = (policy if policy is not None else self.chunk)[axis][0]
            chunk_tabs.append(chunk_factors(lo, dg, hi, rows_per_chunk) if rows_per_chunk else None)
        return dict(line_id=line_id, line_lu=line_lu, line_rows=line_rows, chunk_tabs=chunk_tabs)

    @property
    def launches_per_step(self):
        """kernels launched by one hs2_step"""
        self.ensure_device()
        return int(_cabi.lib().hs2_plan_launches_per_step(self._handle))

    @property
    def x_kernel(self):
        """'whole-line' or 'fold': the kernel a whole-grid hs2_sweep_x of this plan runs"""
        self.ensure_device()
        return ("whole-line", "fold", "tma", "warp")[int(_cabi.lib().hs2_plan_x_kernel(self._handle))]

    def last_kernels(self):
        """names of the kernel variants the last x, y and z sweeps of this plan launched
        (hs2_plan_last_kernel): the parity tests assert the path they mean to cover"""
        self.ensure_device()
        lib = _cabi.lib()
        return tuple(lib.hs2_kernel_name(lib.hs2_plan_last_kernel(self._handle, a)).decode() for a in range(3))

    @property
    def n_unique(self):
        if self._handle is not None:
            return tuple(int(self.axis_info(a).n_unique) for a in range(3))
        return tuple(int(t.shape[0]) for t in self.line_lu)

    # ------------------------------------------------------------- device side
    def ensure_device(self, device=None):
        if self._handle is not None and (device is None or torch.device(device) == self._dev):
            return
        if not torch.cuda.is_available():
            raise RuntimeError("heatsim2_b200 needs a CUDA device (B200, sm_100a); there is no CPU time-step path")
        lib = _cabi.lib()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.release()
        self._dev = dev
        self.d_class_id = self.class_id.to(dev)
        self.class_id = self.d_class_id            # keep a single copy
        if os.environ.get("HS2_TABLES", "native") != "numpy":
            return self._build_native(lib, dev)
        chunk = self.chunk
        self.d_coef = torch.from_numpy(self.scaled_coef).to(dev)
        self.d_line_id = [t.to(dev).contiguous() for t in self.line_id]
        self.d_line_lu = [torch.from_numpy(t).to(dev).contiguous() for t in self.line_lu]
        self.d_chunk = [None if t is None else [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in t]
                        for t in self.chunk_tabs]
        self._utab, self._d_ucode = [None] * 3, [None] * 3
        desc = _cabi.PlanDesc()
        desc.nz, desc.ny, desc.nx = self.shape
        desc.n_classes = self.n_classes
        desc.class_id_bytes = self.d_class_id.element_size()
        desc.d_class_id = self.d_class_id.data_ptr()
        desc.d_class_coef = self.d_coef.data_ptr()
        for a in range(3):
            ax = desc.axis[a]
            ax.d_line_id = self.d_line_id[a].data_ptr()
            ax.n_unique = self.d_line_lu[a].shape[0]
            ax.d_lu = self.d_line_lu[a].data_ptr()
            ax.chunk, ax.n_chunks = chunk[a]
            if self.d_chunk[a] is not None:
                ax.d_tab, ax.d_GE = (t.data_ptr() for t in self.d_chunk[a])
                ax.pitch = self.d_chunk[a][0].shape[2]
                ax.band = interface_band(self.chunk_tabs[a][1])
                if (not self.slab or a != 2) and "xyz"[a] in self.utab_axes:
                    Mc, Pc = chunk[a]
                    utab, ucode = uniform_chunks(self.chunk_tabs[a][0], self.shape[2 - a], Mc, Pc, self.line_id[a])
                    self._utab[a] = np.ascontiguousarray(utab)                    # host, read by hs2_plan_create
                    self._d_ucode[a] = torch.from_numpy(ucode).to(dev)
                    ax.h_utab = self._utab[a].ctypes.data
                    ax.d_ucode = self._d_ucode[a].data_ptr()
                if a == 0 and x_warp_applies(self.shape[2], self.n_classes) and chunk[0] == (16, self.shape[2] // 16) \
                        and self.n_classes <= 64:
                    code, xw, xw_band = ghost_uniform_tables(*self.line_rows[0], 16)
                    self.d_xw_code = torch.from_numpy(code).to(dev)
                    self.d_xw_tab = torch.from_numpy(np.ascontiguousarray(xw)).to(dev)
                    ax.d_xw_code, ax.d_xw_tab, ax.xw_band = self.d_xw_code.data_ptr(), self.d_xw_tab.data_ptr(), xw_band
                if a == 0:
                    M, P = chunk[0]
                    self.d_tab_il = torch.from_numpy(interleave_chunks(self.chunk_tabs[0][0], self.shape[2], M, P)).to(dev)
                    ax.d_tab_il = self.d_tab_il.data_ptr()
        desc.device = dev.index
        desc.flags = self.flags
        if self.slab is not None:
            desc.z_chunk0 = self.slab["k0"] // chunk[2][0]
            desc.z_chunks_global = chunk[2][1]
        handle = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _cabi.check(lib.hs2_plan_create(ctypes.byref(desc), ctypes.byref(handle)))
        self._desc = desc
        self._handle = handle

    def _build_native(self, lib, dev):
        """hs2_plan_build: the library derives every table from the class ids and the class coefficient rows."""
        b = _cabi.BuildDesc()
        b.nz, b.ny, b.nx = self.shape
        b.n_classes = self.n_classes
        b.class_id_bytes = self.d_class_id.element_size()
        b.d_class_id = self.d_class_id.data_ptr()
        b.h_class_coef = self.class_coef.ctypes.data_as(_cabi.c_double_p)
        b.device = dev.index
        b.flags = self.flags
        # rows per chunk: the library's own choice (0) unless the environment forces one
        forced = any(k in os.environ for k in ("HS2_CHUNK", "HS2_CHUNK_X"))
        policy = self.chunk
        for a in range(3):
            b.chunk[a] = (policy[a][0] or -1) if forced else 0
        b.utab_axes = sum(1 << a for a in range(3) if "xyz"[a] in self.utab_axes)
        gid = None
        if self.slab is not None:
            if self._global_ids is None:
                raise RuntimeError("the global class ids of this slab plan were released")
            gid = self._global_ids.to(dev)
            b.d_class_id_global = gid.data_ptr()
            b.nz_global, b.k0 = self.slab["nz_global"], self.slab["k0"]
            b.chunk[2] = policy[2][0]
        handle = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _cabi.check(lib.hs2_plan_build(ctypes.byref(b), ctypes.byref(handle)))
        self._handle = handle
        self._desc = b
        del gid
        if self._np is None:
            self._global_ids = None        # 1 byte per GLOBAL cell: not kept on a rank that never inspects the tables
        got = [(i.chunk, i.n_chunks) for i in (self.axis_info(a) for a in range(3))]
        if not forced and got != [tuple(c) for c in policy]:
            raise AssertionError("chunk policy of plan.py %r and of hs2_plan_build %r differ" % (policy, got))

    def release(self):
        if self._handle is not None:
            _cabi.lib().hs2_plan_destroy(self._handle)
            self._handle = None
        self._bufs = {}

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def _buf(self, name, dtype=torch.float64, shape=None):
        b = self._bufs.get(name)
        shape = self.shape if shape is None else shape
        if b is None or b.shape != torch.Size(shape):
            b = torch.empty(shape, dtype=dtype, device=self._dev)
            self._bufs[name] = b
        return b

    # ------------------------------------------------------------- sources
    def _source(self, t, dt, volumetric_elements, volumetric):
        """Sources of the step at time ``t`` as the C ABI's hs2_source.  Returns
        (Source struct or None, keep-alive list)."""
        table, dense = self.evaluate_sources(t, dt, volumetric_elements, volumetric)
        if table is None and dense is None:
            return None, None
        src = _cabi.Source()
        keep = [table]
        if table is not None:
            src.d_vol_elements = self._vol_elements_dev(volumetric_elements).data_ptr()
            src.h_value = table.ctypes.data_as(_cabi.c_double_p)
        if dense is not None:
            d = torch.from_numpy(np.ascontiguousarray(dense)).to(self._dev)
            keep.append(d)
            src.d_dense = d.data_ptr()
        return src, keep

    @staticmethod
    def source_active(t, volumetric):
        """True when a step taken at time ``t`` may carry a volumetric source (the time conditions of
        ``evaluate_sources`` / alternatingdirection_c_pyx.pyx:301-383, without touching any array)."""
        from . import (NO_SOURCE, STEPPED_SOURCE, IMPULSE_SOURCE, IMPULSE_POINT_SOURCE_JOULES,
                       SPATIALLY_Z_DECAYING_TEMPORAL_IMPULSE)
        for entry in (volumetric if volumetric is not None else ()):
            kind = entry[0]
            if kind == NO_SOURCE:
                continue
            if kind == STEPPED_SOURCE:
                if t >= entry[1] and t <= entry[2]:
                    return True
            elif kind in (IMPULSE_SOURCE, IMPULSE_POINT_SOURCE_JOULES, SPATIALLY_Z_DECAYING_TEMPORAL_IMPULSE):
                if t == entry[1]:    # the three impulse kinds fire at exactly t == t_impulse
                    return True
            else:
                return True          # unknown kind: let evaluate_sources raise for it
        return False

    def evaluate_sources(self, t, dt, volumetric_elements, volumetric):
        """Evaluate the volumetric sources active at time ``t`` (reference
        alternatingdirection_c_pyx.pyx:294-386) on the host.  Returns
        ``(table, dense)``: a 256-entry W/m^3 value per volumetric class (at most
        MAX_TABLE_SOURCES non-zero) and/or a dense [nz,ny,nx] array; None when absent."""
        from . import (IMPULSE_SOURCE, STEPPED_SOURCE, IMPULSE_POINT_SOURCE_JOULES,
                       SPATIALLY_Z_DECAYING_TEMPORAL_IMPULSE, NO_SOURCE)
        table = np.zeros(256)
        dense = None
        for idx, entry in enumerate(volumetric):
            kind = entry[0]
            if kind == NO_SOURCE:
                continue
            if kind == IMPULSE_SOURCE:
                if t == entry[1]:
                    table[idx] = entry[2] / dt
            elif kind == STEPPED_SOURCE:
                if t >= entry[1] and t <= entry[2]:
                    table[idx] = entry[3]
            elif kind == IMPULSE_POINT_SOURCE_JOULES:
                if t == entry[1]:
                    if np.ndim(self.volume_array) > 0:
                        mask = _to_numpy(volumetric_elements) == idx
                        dense = np.zeros(self.shape) if dense is None else dense
                        dense[mask] = entry[2] / (np.asarray(self.volume_array)[mask] * dt)
                    else:
                        table[idx] = entry[2] / (self.volume_array * dt)
            elif kind == SPATIALLY_Z_DECAYING_TEMPORAL_IMPULSE:
                (_, impulse_time, decaydirec, offset, z_ndgrid, decay_dz, joulesperm2, characlength) = entry
                if t == impulse_time:
                    decaydirec = np.asarray(decaydirec, dtype=float)
                    assert (decaydirec == np.array([1.0, 0.0, 0.0])).all() or (decaydirec == np.array([-1.0, 0.0, 0.0])).all()
                    offset = float(offset)
                    decay_dz = abs(float(decay_dz))
                    joulesperm2 = float(joulesperm2)
                    characlength = float(characlength)
                    assert characlength > 0
                    z_ndgrid = np.asarray(z_ndgrid)
                    if self.slab is not None and z_ndgrid.ndim == 3 and z_ndgrid.shape[0] == self.slab["nz_global"] \
                            and z_ndgrid.shape[0] != self.shape[0]:
                        # the tuple describes the global grid; this plan owns planes [k0, k0 + nz)
                        z_ndgrid = z_ndgrid[self.slab["k0"]:self.slab["k0"] + self.shape[0]]
                    centre = z_ndgrid * decaydirec[0] - offset
                    left = centre - decay_dz / 2.0
                    right = centre + decay_dz / 2.0
                    use = (_to_numpy(volumetric_elements) == idx) & (right > 0.0)
                    left = np.where(left < 0.0, 0.0, left)
                    frac = -np.exp(-right[use] / characlength) + np.exp(-left[use] / characlength)
                    dense = np.zeros(self.shape) if dense is None else dense
                    dense[use] = frac * joulesperm2 / (decay_dz * dt)
            else:
                raise ValueError("unknown volumetric source type %r" % (kind,))
        if np.count_nonzero(table) > self.MAX_TABLE_SOURCES:
            # the kernels' table form holds 8 classes; more simultaneously active regions
            # (the reference allows all 256, alternatingdirection_c_pyx.pyx:301-383) go dense
            dense = np.zeros(self.shape) if dense is None else dense
            dense += table[_to_numpy(volumetric_elements)]
            table[:] = 0.0
        return (table if table.any() else None), dense

    def _vol_elements_dev(self, volumetric_elements):
        if isinstance(volumetric_elements, torch.Tensor) and volumetric_elements.is_cuda:
            if volumetric_elements.dtype != torch.uint8 or tuple(volumetric_elements.shape) != self.shape:
                raise ValueError("volumetric_elements must be uint8 with the grid's shape")
            return volumetric_elements.contiguous()
        key = id(volumetric_elements)
        if self._vol_dev is None or self._vol_key != key:
            arr = _to_numpy(volumetric_elements)
            if arr.dtype != np.uint8 or arr.shape != self.shape:
                raise ValueError("volumetric_elements must be uint8 with the grid's shape")
            self._vol_dev = torch.from_numpy(np.ascontiguousarray(arr)).to(self._dev)
            self._vol_key = key
            self._vol_ref = volumetric_elements       # pin the id
        return self._vol_dev

    # ------------------------------------------------------------- stepping
    def step_device(self, T_in, T_out, t, dt, volumetric_elements, volumetric, halo_lo=None, halo_hi=None):
        """One step on device tensors (float64, contiguous, plan's device).
        ``T_out`` may be ``T_in``."""
        self.ensure_device(T_in.device)
        lib = _cabi.lib()
        src, keep = (None, None)
        if volumetric is not None and len(volumetric):
            src, keep = self._source(t, dt, volumetric_elements, volumetric)
        work = self._buf("work")
        stream = torch.cuda.current_stream(self._dev).cuda_stream
        rc = lib.hs2_step(self._handle, T_in.data_ptr(), T_out.data_ptr(), work.data_ptr(),
                          ctypes.byref(src) if src is not None else None,
                          halo_lo.data_ptr() if halo_lo is not None else None,
                          halo_hi.data_ptr() if halo_hi is not None else None,
                          ctypes.c_void_p(stream))
        _cabi.check(rc)
        if keep is not None and len(keep) > 1:
            torch.cuda.current_stream(self._dev).synchronize()    # dense source buffer must outlive the launch
        return T_out

    # ------------------------------------------- many steps per call, observation on the device
    def observe(self, T, cells, probe_out, surface_out, dz):
        """hs2_observe: probe cells (flat int64 indices, device) and / or the insulated z-min surface temperature
        (reference heatsim2/surface_temperature.py:4-36) of the device field ``T`` into device tensors."""
        lib = _cabi.lib()
        stream = ctypes.c_void_p(torch.cuda.current_stream(self._dev).cuda_stream)
        use_p = cells is not None and probe_out is not None
        _cabi.check(lib.hs2_observe(self._handle, T.data_ptr(), cells.data_ptr() if use_p else None,
                                    cells.numel() if use_p else 0, probe_out.data_ptr() if use_p else None,
                                    surface_out.data_ptr() if surface_out is not None else None,
                                    float(dz) if dz is not None else 0.0, stream))

    def run_steps_device(self, T_even, T_odd, first_step, nsteps, every=1, cells=None, probe_rec=None, surf_rec=None,
                         dz=None, first_row=0, use_graph=True):
        """hs2_run_steps: the source-free steps ``first_step .. first_step + nsteps - 1`` in one library call.
        Step n reads ``T_even`` for even n, ``T_odd`` for odd n, and writes the other tensor; after every step with
        ``(n + 1) % every == 0`` the probes / surface estimate go to row ``first_row, first_row + 1, ...`` of the
        record tensors.  Asynchronous on the current stream."""
        self.ensure_device(T_even.device)
        for T in (T_even, T_odd):
            self._check_field(T)
            if T.dtype != torch.float64 or not T.is_contiguous() or T.device != T_even.device:
                raise ValueError("run_steps_device: fields must be contiguous float64 tensors on one device")
        lib = _cabi.lib()
        work = self._buf("work")
        counter = self._bufs.get("row_counter")
        if counter is None:
            counter = self._bufs["row_counter"] = torch.zeros(1, dtype=torch.int64, device=self._dev)
        use_p = cells is not None and probe_rec is not None
        if use_p or surf_rec is not None:
            counter.fill_(int(first_row))
        stream = ctypes.c_void_p(torch.cuda.current_stream(self._dev).cuda_stream)
        _cabi.check(lib.hs2_run_steps(self._handle, T_even.data_ptr(), T_odd.data_ptr(), work.data_ptr(), int(first_step),
                                      int(nsteps), int(every), cells.data_ptr() if use_p else None,
                                      cells.numel() if use_p else 0, probe_rec.data_ptr() if use_p else None,
                                      surf_rec.data_ptr() if surf_rec is not None else None,
                                      float(dz) if dz is not None else 0.0, counter.data_ptr(), 1 if use_graph else 0, stream))

    def upload(self, array, dev):
        """numpy array (or host tensor) -> device tensor ``dev`` at full PCIe speed: page-locked sources by one DMA,
        pageable ones through the pinned staging chunks."""
        arr = np.ascontiguousarray(_to_numpy(array), dtype=np.float64)
        self._check_field(arr)
        h = torch.from_numpy(arr)
        if h.is_pinned():
            dev.copy_(h, non_blocking=True)
            torch.cuda.current_stream(self._dev).synchronize()
        else:
            self._upload(h, dev)
        return dev

    def download(self, dev):
        """device tensor -> new numpy array (page-locked unless HS2_PINNED_RESULTS=0, like ``run_step``'s results)"""
        if os.environ.get("HS2_PINNED_RESULTS", "1") != "0":
            result = torch.empty(tuple(dev.shape), dtype=torch.float64, pin_memory=True)
            result.copy_(dev, non_blocking=True)
            torch.cuda.current_stream(self._dev).synchronize()
        else:
            result = torch.empty(tuple(dev.shape), dtype=torch.float64)
            self._download(dev, result)
        return result.numpy()

    def _check_field(self, T):
        if tuple(T.shape) != self.shape:
            raise ValueError("Tarray has shape %r, the plan was set up for %r" % (tuple(T.shape), self.shape))

    @staticmethod
    def check_out(out, like):
        """``out=`` goes to the kernels as a raw pointer: it has to look exactly like the input"""
        if not isinstance(out, torch.Tensor):
            raise TypeError("out must be a torch tensor")
        if out.shape != like.shape or out.dtype != like.dtype or out.device != like.device or not out.is_contiguous():
            raise ValueError("out must be a contiguous %s tensor of shape %r on %s (got %s %r on %s, contiguous=%s)"
                             % (like.dtype, tuple(like.shape), like.device, out.dtype, tuple(out.shape), out.device,
                                out.is_contiguous()))

    def timed_sweeps(self, T_in, T_out, events):
        """One source-free step as three separate C-ABI calls with CUDA events
        recorded on the launching stream between them (bench.py's per-kernel
        roofline).  ``events`` = 4 torch.cuda.Event(enable_timing=True)."""
        self.ensure_device(T_in.device)
        lib = _cabi.lib()
        work = self._buf("work")
        ts = torch.cuda.current_stream(self._dev)
        stream = ctypes.c_void_p(ts.cuda_stream)
        events[0].record(ts)
        _cabi.check(lib.hs2_sweep_x(self._handle, T_in.data_ptr(), work.data_ptr(), None, None, None, stream))
        events[1].record(ts)
        _cabi.check(lib.hs2_sweep_y(self._handle, work.data_ptr(), stream))
        events[2].record(ts)
        _cabi.check(lib.hs2_sweep_z(self._handle, T_in.data_ptr(), T_out.data_ptr(), work.data_ptr(), stream))
        events[3].record(ts)

    def run_step(self, t, dt, Tarray, volumetric_elements, volumetric, out=None):
        if isinstance(Tarray, torch.Tensor) and not Tarray.is_cuda:
            # host tensor (ideally pinned) in -> host tensor out, no staging copies
            self._check_field(Tarray)
            if Tarray.dtype != torch.float64 or not Tarray.is_contiguous():
                raise ValueError("Tarray must be a contiguous float64 tensor")
            self.ensure_device()
            with torch.cuda.device(self._dev):
                d_T = self._buf("T")
                d_T.copy_(Tarray, non_blocking=True)
                self.step_device(d_T, d_T, t, dt, volumetric_elements, volumetric)
                if out is None:
                    out = torch.empty(self.shape, dtype=torch.float64).pin_memory()
                else:
                    self.check_out(out, Tarray)
                out.copy_(d_T, non_blocking=True)
                torch.cuda.current_stream(self._dev).synchronize()
            return out
        if isinstance(Tarray, torch.Tensor) and Tarray.is_cuda:
            self._check_field(Tarray)
            if Tarray.dtype != torch.float64:
                raise ValueError("Tarray must be float64")
            T_in = Tarray.contiguous()
            if out is not None:
                self.check_out(out, T_in)
            with torch.cuda.device(T_in.device):
                T_out = torch.empty_like(T_in) if out is None else out
                return self.step_device(T_in, T_out, t, dt, volumetric_elements, volumetric)
        arr = np.ascontiguousarray(_to_numpy(Tarray), dtype=np.float64)
        self._check_field(arr)
        self.ensure_device()
        with torch.cuda.device(self._dev):
            d_T = self._buf("T")
            h_in = torch.from_numpy(arr)
            if h_in.is_pinned():
                d_T.copy_(h_in, non_blocking=True)        # an array this function returned earlier: DMA straight from it
            else:
                self._upload(h_in, d_T)                   # the user's pageable array: staged through pinned chunks
            self.step_device(d_T, d_T, t, dt, volumetric_elements, volumetric)
            # The new array the caller gets (like the reference) lives in page-locked memory (torch's caching host
            # allocator re-uses the block of an array the caller has dropped), so the download is one DMA and, in the
            # usual loop T = run_adi_steps(..., T, ...), so is the next step's upload.  HS2_PINNED_RESULTS=0: plain
            # pageable arrays (staged copies both ways).
            if os.environ.get("HS2_PINNED_RESULTS", "1") != "0":
                result = torch.empty(self.shape, dtype=torch.float64, pin_memory=True)
                result.copy_(d_T, non_blocking=True)
                torch.cuda.current_stream(self._dev).synchronize()
            else:
                result = torch.empty(self.shape, dtype=torch.float64)
                self._download(d_T, result)
            return result.numpy()

    # numpy in / numpy out (the reference's contract): the user's arrays are pageable, so they travel through two
    # pinned staging chunks; the host memcpy of chunk i+1 (torch's multi-threaded copy) overlaps the DMA of chunk i
    STAGE_BYTES = 64 << 20

    def _staging(self):
        b = self._bufs.get("stage")
        if b is None:
            b = [torch.empty(self.STAGE_BYTES // 8, dtype=torch.float64).pin_memory() for _ in range(2)]
            self._bufs["stage"] = b
            self._bufs["stage_ev"] = [torch.cuda.Event() for _ in range(2)]
        return b, self._bufs["stage_ev"]

    def _upload(self, host, dev):
        """pageable host tensor -> device tensor, on the current stream"""
        stage, ev = self._staging()
        src, dst = host.reshape(-1), dev.reshape(-1)
        n, step = src.numel(), stage[0].numel()
        stream = torch.cuda.current_stream(self._dev)
        for c, o in enumerate(range(0, n, step)):
            m = min(step, n - o)
            b = c & 1
            ev[b].synchronize()                       # the DMA that last read this staging buffer is done
            stage[b][:m].copy_(src[o:o + m])
            dst[o:o + m].copy_(stage[b][:m], non_blocking=True)
            ev[b].record(stream)

    def _download(self, dev, host):
        """device tensor -> pageable host tensor; returns when the data is there"""
        stage, ev = self._staging()
        src, dst = dev.reshape(-1), host.reshape(-1)
        n, step = src.numel(), stage[0].numel()
        stream = torch.cuda.current_stream(self._dev)
        ev[0].synchronize()
        ev[1].synchronize()
        pending = None
        for c, o in enumerate(range(0, n, step)):
            m = min(step, n - o)
            b = c & 1
            stage[b][:m].copy_(src[o:o + m], non_blocking=True)
            ev[b].record(stream)
            if pending is not None:
                pb, po, pm = pending
                ev[pb].synchronize()
                dst[po:po + pm].copy_(stage[pb][:pm])
            pending = (b, o, m)
        if pending is not None:
            pb, po, pm = pending
            ev[pb].synchronize()
            dst[po:po + pm].copy_(stage[pb][:pm])

    # ------------------------------------------- host inspection (small grids)
    def reference_matrices(self, stepnum):
        """Materialise stage ``stepnum``'s A (n x 3), B, C (scipy CSR) and D in
        the reference's permuted row order (alternatingdirection_c.c:114) from
        the class tables - for inspection and tests on small grids."""
        import scipy.sparse
        from .alternatingdirection_c_pyx import _STAGES
        if self.n > 4_000_000:
            raise MemoryError("reference_matrices is meant for small grids")
        perm = _STAGES[stepnum][1]
        nz, ny, nx = self.shape
        cid = self.class_id.cpu().numpy().astype(np.int64) & (0xFFFF if self.class_id.dtype == torch.int16 else 0x7FFFFFFF)
        cc = self.class_coef[cid]                       # [nz,ny,nx,8]
        M = cc[..., M_]
        g = {2: (cc[..., GXM], cc[..., GXP]), 1: (cc[..., GYM], cc[..., GYP]), 0: (cc[..., GZM], cc[..., GZP])}
        pshape = [self.shape[a] for a in perm]
        idx = np.arange(self.n).reshape(pshape).transpose(np.argsort(perm))    # permuted flat index of cell (k,j,i)
        sweep = perm[2]
        # weight of each direction in B for this stage: 1 (not yet implicit) or 1/2
        half = {2: 0.5, 1: 0.5 if stepnum >= 1 else 1.0, 0: 0.5 if stepnum >= 2 else 1.0}
        A = np.zeros((self.n, 3))
        A[idx.ravel(), 0] = (-0.5 * g[sweep][0]).ravel()
        A[idx.ravel(), 2] = (-0.5 * g[sweep][1]).ravel()
        A[idx.ravel(), 1] = (M + 0.5 * (g[sweep][0] + g[sweep][1])).ravel()
        rows, cols, vals = [idx.ravel()], [idx.ravel()], [(M - sum(half[a] * (g[a][0] + g[a][1]) for a in range(3))).ravel()]
        crow = {c: [[], [], []] for c in range(stepnum)}
        axis_of_stage = {0: 2, 1: 1, 2: 0}
        for c in range(stepnum):
            a = axis_of_stage[c]
            crow[c][0].append(idx.ravel()); crow[c][1].append(idx.ravel())
            crow[c][2].append((-0.5 * (g[a][0] + g[a][1])).ravel())
        for a in range(3):
            for side, sgn in ((0, -1), (1, 1)):
                gv = g[a][side]
                sl_c = [slice(None)] * 3
                sl_n = [slice(None)] * 3
                sl_c[a] = slice(1, None) if sgn < 0 else slice(None, -1)
                sl_n[a] = slice(None, -1) if sgn < 0 else slice(1, None)
                r = idx[tuple(sl_c)].ravel()
                cn = idx[tuple(sl_n)].ravel()
                v = gv[tuple(sl_c)].ravel()
                keep = v != 0.0
                rows.append(r[keep]); cols.append(cn[keep]); vals.append(half[a] * v[keep])
                for c in range(stepnum):
                    if axis_of_stage[c] == a:
                        crow[c][0].append(r[keep]); crow[c][1].append(cn[keep]); crow[c][2].append(0.5 * v[keep])
        n = self.n
        B = scipy.sparse.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
        C = [scipy.sparse.coo_matrix((np.concatenate(crow[c][2]), (np.concatenate(crow[c][0]), np.concatenate(crow[c][1]))),
                                     shape=(n, n)).tocsr() for c in range(stepnum)]
        D = np.zeros(n)
        D[idx.ravel()] = cc[..., D_].ravel()
        return {"A": A, "B": B, "C": C, "D": D}


def _to_numpy(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def _unique_lines(cid, sub_lut, axis, n_sub):
    """Group the lines of sweep ``axis`` by their sequence of row signatures.

    cid: [nz,ny,nx] integer tensor of class ids; sub_lut: class -> row-signature
    id.  Returns (line_id int32 [n_lines], reps int64 numpy [n_unique, L] of
    signature ids).  Lines are hashed (64-bit polynomial) in z-chunks to bound
    temporary memory, grouped by hash, and the grouping is then verified
    exactly; a hash collision falls back to an exact row-wise unique."""
    nz, ny, nx = cid.shape
    dev = cid.device
    L = (nx, ny, nz)[axis]
    n_lines = cid.numel() // L
    mult = 0x9E3779B97F4A7C15 - (1 << 64)          # odd 64-bit multiplier as int64
    pw = torch.empty(L, dtype=torch.int64)
    acc = 1
    for r in range(L):
        pw[r] = acc if acc < (1 << 63) else acc - (1 << 64)
        acc = (acc * (mult & ((1 << 64) - 1)) + 0x632BE59BD9B4E019) & ((1 << 64) - 1)
    pw = pw.to(dev)
    chunk = max(1, int(32e6 // max(1, ny * nx)))
    if axis == 2:
        h = torch.zeros(ny * nx, dtype=torch.int64, device=dev)
    else:
        h = torch.empty(n_lines, dtype=torch.int64, device=dev)
    for k0 in range(0, nz, chunk):
        k1 = min(nz, k0 + chunk)
        sub = sub_lut[cid[k0:k1].reshape(-1).long()].reshape(k1 - k0, ny, nx) + 1
        if axis == 0:
            h[k0 * ny:k1 * ny] = (sub * pw.view(1, 1, nx)).sum(dim=2).reshape(-1)
        elif axis == 1:
            h[k0 * nx:k1 * nx] = (sub * pw.view(1, ny, 1)).sum(dim=1).reshape(-1)
        else:
            h += (sub * pw[k0:k1].view(-1, 1, 1)).sum(dim=0).reshape(-1)
    uh, inv = torch.unique(h, return_inverse=True)
    nu = uh.numel()
    # first line of every hash group
    first = torch.full((nu,), n_lines, dtype=torch.int64, device=dev)
    first.scatter_reduce_(0, inv, torch.arange(n_lines, device=dev), reduce="amin")
    # order groups by first appearance so ids are deterministic
    order = torch.argsort(first)
    rank = torch.empty_like(order)
    rank[order] = torch.arange(nu, device=dev)
    line_id = rank[inv]
    first = first[order]
    sub_lines = _line_view(sub_lut[cid.reshape(-1).long()].reshape(nz, ny, nx).to(torch.int16 if n_sub < 32768 else torch.int32), axis)
    reps = sub_lines[first]                                  # [nu, L]
    # exact verification, chunked over lines
    ok = True
    step = max(1, int(64e6 // L))
    for l0 in range(0, n_lines, step):
        l1 = min(n_lines, l0 + step)
        if not torch.equal(sub_lines[l0:l1], reps[line_id[l0:l1]]):
            ok = False
            break
    if not ok:
        reps, line_id = torch.unique(sub_lines, dim=0, return_inverse=True)
    return line_id.to(torch.int32).contiguous(), reps.cpu().numpy().astype(np.int64)


class CellwiseBuilder(object):
    """Accumulates per-cell equations given one at a time through the
    reference-style ``add_equation_to_adi_matrices`` and turns them into class
    tables (small problems / custom assemblies)."""

    def __init__(self, shape):
        self.shape = tuple(shape)
        self.class_of_key = {}
        self.coefs = []
        self.class_id = np.full(self.shape, -1, dtype=np.int64)

    def add(self, k, j, i, key, eqdicts):
        from .alternatingdirection_c_pyx import class_coefficients
        c = self.class_of_key.get(key)
        if c is None:
            M, g, D = class_coefficients(eqdicts)
            c = self.class_of_key[key] = len(self.coefs)
            self.coefs.append((M,) + tuple(g) + (D,))
        self.class_id[k, j, i] = c

    def add_stage(self, stepnum, pos, eqdict):
        """one stage's dictionary of one cell (pyadi_step.add_equation); the cell is classified as soon as all three
        stages are there"""
        pending = self.__dict__.setdefault("_pending", {})
        entry = pending.setdefault(pos, [None, None, None])
        entry[stepnum] = eqdict
        if all(d is not None for d in entry):
            key = tuple(tuple(sorted(d.items())) for d in entry)
            self.add(pos[0], pos[1], pos[2], key, entry)
            del pending[pos]

    def finish(self, dt, volume_array):
        if getattr(self, "_pending", None):
            raise ValueError("%d cells have equations for some stages only" % len(self._pending))
        if (self.class_id < 0).any():
            raise ValueError("some cells have no equation")
        return AdiPlan(self.shape, self.class_id, np.array(self.coefs), dt, volume_array)

"""Thread-level numpy transcription of csrc/kernels_xm.cu (the z-marching x
sweep): same index formulas, shared-memory layout, copy-box geometry, slot
rotation and mbarrier bookkeeping as the CUDA kernel, executed one thread at a
time.  TEST INFRASTRUCTURE - it lets the CPU suite check the kernel's indexing
and its load/wait protocol (every wait has exactly one completed copy to
consume, every slice holds the plane the code believes it holds) against the
plain stencil + line solve of tests/emul.py.  Not part of the product."""
import numpy as np

from heatsim2_b200.plan import T_INV, T_F, T_C, T_S, T_CP

SLOTS, PAD, TP = 4, 2, 2


def row_pitch(P, M):
    s = P * (M + PAD)
    while (s & 3) != 2:
        s += 2
    return s


def geometry(nz, ny, nx, M, P, KR, R=8):
    RW = R + 2
    if R * P > 256 or nx % 2 or ny < RW:
        return None
    cpt = nx // 2
    if 2 * cpt > 512:
        return None
    RPT = 2 if 4 * cpt <= 512 else 4
    threads = max((R // RPT) * cpt, R * P)
    threads = -(-threads // 32) * 32
    if threads > 512:
        return None
    g = dict(threads=threads, RPT=RPT, solvers=R * P, R=R)
    g["BX"] = nx if nx <= 256 else 256
    g["NXB"] = -(-nx // g["BX"])
    g["box_stride"] = -(-(RW * g["BX"]) // 16) * 16
    g["slot_stride"] = g["NXB"] * g["box_stride"]
    g["Sr"] = row_pitch(P, M)
    g["buf_alias"] = R * g["Sr"] <= g["slot_stride"]
    g["KR"] = max(1, min(KR, nz))
    g["tiles_y"] = -(-ny // R)
    g["n_items"] = -(-nz // g["KR"]) * g["tiles_y"]
    return g


def sweep_x(plan, T, KR=32, n_blocks=3, R=8):
    """d1 of stage 0 for the whole grid, computed the way sweep_xm_kernel does."""
    RW = R + 2
    nz, ny, nx = plan.shape
    M, P = plan.chunk[0]
    g = geometry(nz, ny, nx, M, P, KR, R)
    assert g is not None, "grid outside the kernel's range"
    BX, NXB, box_stride, slot_stride, Sr = g["BX"], g["NXB"], g["box_stride"], g["slot_stride"], g["Sr"]
    nthreads, RPT, nsolvers, alias = g["threads"], g["RPT"], g["solvers"], g["buf_alias"]
    RG, cpt = R // RPT, nx // 2
    tab_u, GE_u = plan.chunk_tabs[0]          # [nu, 5, pitch], [nu, P, 2P]
    pitch = tab_u.shape[2]
    line_id = plan.line_id[0].cpu().numpy()   # x-lines k*ny + j
    cid = plan.class_id.cpu().numpy().astype(np.int64) & 0xFFFF
    coef = plan.scaled_coef                   # [nc, 8]: gx-,gx+,gy-,gy+,gz-,gz+,D,M
    band = P                                  # full rows of the interface operator (superset of the device band)
    W = np.full((nz, ny, nx), np.nan)

    for block in range(min(n_blocks, g["n_items"])):
        slots = np.full(SLOTS * slot_stride, np.nan)
        tag = [None] * SLOTS                  # (plane, j0) a slice holds
        completed = [0] * SLOTS
        waited = [0] * SLOTS
        phase_bits = 0
        buf_own = np.full(R * Sr, np.nan)
        Y = np.zeros(2 * P * R)
        Es = np.zeros(P * R)
        # shared-memory tables of the block's first line
        kr0, jt0 = block // g["tiles_y"], block % g["tiles_y"]
        lid_c = line_id[(kr0 * g["KR"]) * ny + jt0 * R]
        per_plane = P * (M + TP)
        s_tab = np.zeros(5 * per_plane)
        for e in range(5 * per_plane):
            pl, rem = divmod(e, per_plane)
            pp, t = divmod(rem, M + TP)
            row = pp * M + t
            s_tab[e] = tab_u[lid_c, pl, row] if (t < M and row < pitch) else 0.0
        s_ge = np.zeros(P * (2 * P + 2))
        for e in range(P * (2 * P + 2)):
            pp, q = divmod(e, 2 * P + 2)
            s_ge[e] = GE_u[lid_c, pp, q] if q < 2 * P else 0.0

        def tma_fill(slot, j0, plane_k):
            base = slot * slot_stride
            for xb in range(NXB):
                for q in range(RW):
                    j = j0 - 1 + q
                    for c in range(BX):
                        i = xb * BX + c
                        ok = 0 <= j < ny and i < nx
                        slots[base + xb * box_stride + q * BX + c] = T[plane_k, j, i] if ok else 0.0

        for item in range(block, g["n_items"], n_blocks):
            kr, jt = divmod(item, g["tiles_y"])
            j0 = jt * R
            ka = kr * g["KR"]
            kb = min(nz, ka + g["KR"])
            n_it = kb - ka
            nrows = min(R, ny - j0)

            def needed(q):
                pl = ka - 1 + q
                return 0 <= pl < nz and q <= n_it + 1

            def issue(q):
                s = q & (SLOTS - 1)
                assert completed[s] == waited[s], "copy issued into a slice whose previous copy was never consumed"
                tma_fill(s, j0, ka - 1 + q)
                tag[s] = (ka - 1 + q, j0)
                completed[s] += 1

            def wait_slot(q):
                nonlocal phase_bits
                s = q & (SLOTS - 1)
                parity = (phase_bits >> s) & 1
                assert completed[s] == waited[s] + 1, "wait without exactly one completed copy"
                assert (completed[s] & 1) != parity, "parity the kernel waits on would never flip"
                waited[s] += 1
                phase_bits ^= 1 << s

            for q in range(3):
                if needed(q):
                    issue(q)
            for it in range(n_it):
                k = ka + it
                if needed(it + 3):
                    issue(it + 3)
                if it == 0:
                    if needed(0):
                        wait_slot(0)
                    wait_slot(1)
                if k + 1 < nz:
                    wait_slot(it + 2)
                sC = (it + 1) & (SLOTS - 1)
                sL = (it & (SLOTS - 1)) if k > 0 else sC
                sH = ((it + 2) & (SLOTS - 1)) if k + 1 < nz else sC
                assert tag[sC] == (k, j0) and tag[sL] == (max(k - 1, 0), j0) and tag[sH] == (min(k + 1, nz - 1), j0)
                Cn, Lo, Hi = sC * slot_stride, sL * slot_stride, sH * slot_stride
                # solve buffer: own region, or the slice of plane k-1 (slot it & 3)
                if alias:
                    bs = it & (SLOTS - 1)
                    assert completed[bs] == waited[bs], "solve buffer placed in a slice with a copy in flight"
                    buf, boff = slots, bs * slot_stride
                else:
                    buf, boff = buf_own, 0
                # ---------------- phase 1: all reads, barrier, then the stores
                outv = {}
                for tid in range(nthreads):
                    rg = tid // cpt
                    if rg >= RG:
                        continue
                    i = 2 * (tid - rg * cpt)
                    r0 = rg * RPT
                    so = (i // BX) * box_stride + (i % BX)
                    rq = []
                    for q in range(RPT + 2):
                        j = min(max(j0 + r0 + q - 1, 0), ny - 1)
                        rq.append((j - (j0 - 1)) * BX)
                    for r in range(RPT):
                        jj = min(j0 + r0 + r, ny - 1)
                        for e in range(2):
                            cf = coef[cid[k, jj, i + e]]
                            tt = slots[Cn + so + rq[r + 1] + e]
                            vym = slots[Cn + so + rq[r] + e]
                            vyp = slots[Cn + so + rq[r + 2] + e]
                            vzm = slots[Lo + so + (r0 + r + 1) * BX + e]
                            vzp = slots[Hi + so + (r0 + r + 1) * BX + e]
                            rr = cf[2] * (vym - tt) + cf[3] * (vyp - tt) + cf[4] * (vzm - tt) + cf[5] * (vzp - tt)
                            outv[(tid, r, e)] = 2.0 * tt + rr
                if alias:
                    tag[bs] = None                        # the slice of plane k-1 is gone from here on
                for (tid, r, e), val in outv.items():
                    rg = tid // cpt
                    i = 2 * (tid - rg * cpt)
                    bo = i + PAD * (i // M)
                    buf[boff + bo + (rg * RPT + r) * Sr + e] = val
                # ---------------- phase 2 (forward)
                v_all = {}
                for tid in range(nsolvers):
                    r2, p2 = tid % R, tid // R
                    pc = p2
                    c0 = pc * M
                    rows = min(M, nx - c0)
                    lid = line_id[k * ny + j0 + (r2 if r2 < nrows else 0)]
                    mine = boff + r2 * Sr + pc * (M + PAD)
                    tab_s = lid == lid_c

                    def tb(pl, t, tab_s=tab_s, pc=pc, c0=c0, lid=lid):
                        if tab_s:
                            return s_tab[pl * per_plane + pc * (M + TP) + t]
                        return tab_u[lid, pl, c0 + t]
                    v = [buf[mine + t] if t < rows else 0.0 for t in range(M)]
                    prev, yf = 0.0, 0.0
                    for t in range(rows):
                        prev = v[t] * tb(T_INV, t) - tb(T_F, t) * prev
                        v[t] = prev
                        yf += tb(T_C, t) * prev
                    Y[(2 * p2) * R + r2] = yf
                    Y[(2 * p2 + 1) * R + r2] = prev
                    v_all[tid] = (v, tb, rows, lid, tab_s, pc, r2, p2, mine)
                E_all = {}
                for tid in range(nsolvers):
                    v, tb, rows, lid, tab_s, pc, r2, p2, mine = v_all[tid]
                    E = 0.0
                    for q in range(max(0, pc - band), min(P - 1, pc + band) + 1):
                        if tab_s:
                            g0, g1 = s_ge[pc * (2 * P + 2) + 2 * q], s_ge[pc * (2 * P + 2) + 2 * q + 1]
                        else:
                            g0, g1 = GE_u[lid, pc, 2 * q], GE_u[lid, pc, 2 * q + 1]
                        E += g0 * Y[(2 * q) * R + r2] + g1 * Y[(2 * q + 1) * R + r2]
                    E_all[tid] = E
                    Es[p2 * R + r2] = E
                for tid in range(nsolvers):
                    v, tb, rows, lid, tab_s, pc, r2, p2, mine = v_all[tid]
                    E = E_all[tid]
                    alpha = Es[(p2 - 1) * R + r2] if p2 > 0 else 0.0
                    nxt = E
                    for t in range(rows - 1, -1, -1):
                        if t < rows - 1:
                            nxt = (v[t] - alpha * tb(T_S, t)) - tb(T_CP, t) * nxt
                        v[t] = nxt
                    for t in range(rows):
                        buf[mine + t] = v[t]
                # ---------------- phase 3
                for tid in range(nthreads):
                    rg = tid // cpt
                    if rg >= RG:
                        continue
                    i = 2 * (tid - rg * cpt)
                    r0 = rg * RPT
                    so = (i // BX) * box_stride + (i % BX)
                    bo = i + PAD * (i // M)
                    for r in range(RPT):
                        if r0 + r >= nrows:
                            continue
                        for e in range(2):
                            t0 = slots[Cn + so + (r0 + r + 1) * BX + e]
                            w = buf[boff + bo + (r0 + r) * Sr + e]
                            assert np.isnan(W[k, j0 + r0 + r, i + e]), "cell written twice"
                            W[k, j0 + r0 + r, i + e] = w - 2.0 * t0
            assert completed == waited, "copies left unconsumed at the end of an item"
    assert not np.isnan(W).any(), "cells never written"
    return W

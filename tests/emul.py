"""numpy transcription of the CUDA kernels' arithmetic, driven by the same plan
tables.  TEST INFRASTRUCTURE: lets the host-side logic of the multi-GPU path
(slab partition, halo exchange, all-gather layout, global z-line tables) run
under gloo on CPU ranks.  Not part of the product; the product has no CPU path.
"""
import numpy as np
import torch

from heatsim2_b200 import dist as hdist
from heatsim2_b200.plan import T_INV, T_F, T_C, T_S, T_CP


def _lines(a3, axis):
    nz, ny, nx = a3.shape
    if axis == 0:
        return a3.reshape(nz * ny, nx)
    if axis == 1:
        return a3.transpose(0, 2, 1).reshape(nz * nx, ny)
    return a3.transpose(1, 2, 0).reshape(ny * nx, nz)


def _unlines(a2, shape, axis):
    nz, ny, nx = shape
    if axis == 0:
        return a2.reshape(nz, ny, nx)
    if axis == 1:
        return a2.reshape(nz, nx, ny).transpose(0, 2, 1)
    return a2.reshape(ny, nx, nz).transpose(2, 0, 1)


def chunk_forward(d, tab, M, row0=0):
    """d [n, L_local]; tab [n, 5, pitch] already gathered per line.  Returns
    (u, Y[n, 2P]) with Y interleaved (yf_0, yl_0, ...)."""
    n, L = d.shape
    P = -(-L // M)
    u = np.zeros_like(d)
    Y = np.zeros((n, 2 * P))
    for p in range(P):
        prev = np.zeros(n)
        acc = np.zeros(n)
        for r in range(p * M, min(L, (p + 1) * M)):
            g = row0 + r
            prev = d[:, r] * tab[:, T_INV, g] - tab[:, T_F, g] * prev
            u[:, r] = prev
            acc += tab[:, T_C, g] * prev
        Y[:, 2 * p], Y[:, 2 * p + 1] = acc, prev
    return u, Y


def chunk_backward(u, tab, M, E, alpha, row0=0):
    """E, alpha [n, P] for the local chunks."""
    n, L = u.shape
    P = -(-L // M)
    x = np.zeros_like(u)
    for p in range(P):
        r1 = min(L, (p + 1) * M)
        nxt = E[:, p].copy()
        x[:, r1 - 1] = nxt
        for r in range(r1 - 2, p * M - 1, -1):
            g = row0 + r
            nxt = (u[:, r] - alpha[:, p] * tab[:, T_S, g]) - tab[:, T_CP, g] * nxt
            x[:, r] = nxt
    return x


def solve_axis(plan, W, axis):
    """single-device chunked solve of all lines of `axis` (kernels_strided.cu / kernels_xf.cu phase 2)"""
    M, P = plan.chunk[axis]
    tab_u, GE_u = plan.chunk_tabs[axis]
    lid = plan.line_id[axis].cpu().numpy()
    d = np.ascontiguousarray(_lines(W, axis))
    tab = tab_u[lid]
    u, Y = chunk_forward(d, tab, M)
    E = np.einsum("npq,nq->np", GE_u[lid], Y)
    alpha = np.concatenate([np.zeros((len(lid), 1)), E[:, :-1]], axis=1)
    x = chunk_backward(u, tab, M, E, alpha)
    return np.ascontiguousarray(_unlines(x, W.shape, axis))


def rhs(plan, T, src_dense, halo_lo, halo_hi):
    """stage-0 right hand side (unfolded form; kernels_xf.cu folds the x-term into the solve)"""
    cid = plan.class_id.cpu().numpy().astype(np.int64) & 0xFFFF
    c = plan.scaled_coef[cid]
    Tp = np.pad(T, 1, mode="edge")
    if halo_lo is not None:
        Tp[0, 1:-1, 1:-1] = halo_lo
    if halo_hi is not None:
        Tp[-1, 1:-1, 1:-1] = halo_hi
    tc = Tp[1:-1, 1:-1, 1:-1]
    r = c[..., 0] * (Tp[1:-1, 1:-1, :-2] - tc) + c[..., 1] * (Tp[1:-1, 1:-1, 2:] - tc)
    r += c[..., 2] * (Tp[1:-1, :-2, 1:-1] - tc) + c[..., 3] * (Tp[1:-1, 2:, 1:-1] - tc)
    r += c[..., 4] * (Tp[:-2, 1:-1, 1:-1] - tc) + c[..., 5] * (Tp[2:, 1:-1, 1:-1] - tc)
    if src_dense is not None:
        r += c[..., 6] * src_dense
    return r


def source_array(plan, src, keep):
    """dense W/m^3 array from the Source struct the plan built"""
    if src is None:
        return None
    out = np.zeros(plan.shape)
    if src.h_value:
        table = np.ctypeslib.as_array(src.h_value, (256,))
        out += table[plan._vol_dev.cpu().numpy()]
    if len(keep) > 1:
        out += keep[1].cpu().numpy()
    return out


def step_single(plan, T, src_dense=None):
    """whole single-device step, emulated"""
    W = rhs(plan, T, src_dense, None, None)
    W = solve_axis(plan, W, 0)
    W = solve_axis(plan, W, 1)
    W = solve_axis(plan, W, 2)
    return T + W


class EmulDistPlan(hdist.DistPlan):
    """DistPlan whose four kernels are the numpy transcriptions above (CPU
    tensors, gloo)."""

    def _prepare(self, T_in):
        pass

    def _stream(self, t):
        return None

    def _k_sweep_x(self, T_in, work, src, keep, halo_lo, halo_hi):
        W = rhs(self.plan, T_in.numpy(), source_array(self.plan, src, keep),
                None if halo_lo is None else halo_lo.numpy(), None if halo_hi is None else halo_hi.numpy())
        work.copy_(torch.from_numpy(solve_axis(self.plan, W, 0)))

    def _k_sweep_y(self, work):
        work.copy_(torch.from_numpy(solve_axis(self.plan, work.numpy().copy(), 1)))

    def _z_tables(self):
        plan = self.plan
        tab_u, GE_u = plan.chunk_tabs[2]
        lid = plan.line_id[2].cpu().numpy()
        return tab_u[lid], GE_u[lid]

    def _k_z_forward(self, work, Y, line0, n_lines):
        tab, _ = self._z_tables()
        sl = slice(line0, line0 + n_lines)
        full = np.ascontiguousarray(_lines(work.numpy(), 2))
        u, Yl = chunk_forward(full[sl], tab[sl], self.chunk, row0=self.k0)      # u is not kept (recomputed later)
        Y.copy_(torch.from_numpy(np.ascontiguousarray(Yl.T)))           # [2P_loc, n_lines]

    def _k_z_backward(self, T_in, T_out, work, Yall, line0, n_lines):
        tab, GE = self._z_tables()
        sl = slice(line0, line0 + n_lines)
        # slabs outside the exchange radius were never received: their slots hold
        # garbage that must not matter (the operator is banded) - poison them
        Yg = Yall.numpy().T.copy()                                        # [n_lines, 2P_glob]
        rows = 2 * self.p_loc
        for r in range(self.world):
            if abs(r - self.rank) > self.hops and 2 * self.hops < self.world - 1:
                Yg[:, r * rows:(r + 1) * rows] = 0.0
        P_glob = GE.shape[1]
        band = self.band
        mask = (np.abs(np.arange(P_glob)[:, None] - np.arange(P_glob)[None, :]) <= band).astype(float)
        mask2 = np.repeat(mask, 2, axis=1)                                # interleaved (yf, yl) columns
        Eg = np.einsum("npq,nq->np", GE[sl] * mask2[None], Yg)
        c0 = self.k0 // self.chunk
        E = Eg[:, c0:c0 + self.p_loc]
        alpha = np.concatenate([np.zeros((E.shape[0], 1)), Eg], axis=1)[:, c0:c0 + self.p_loc]
        full_d = np.ascontiguousarray(_lines(work.numpy(), 2))
        u, _ = chunk_forward(full_d[sl], tab[sl], self.chunk, row0=self.k0)
        x = chunk_backward(u, tab[sl], self.chunk, E, alpha, row0=self.k0)
        out = np.ascontiguousarray(_lines(T_out.numpy(), 2)) if T_out is not T_in else np.ascontiguousarray(_lines(T_in.numpy(), 2))
        tin = np.ascontiguousarray(_lines(T_in.numpy(), 2))
        out[sl] = tin[sl] + x
        T_out.copy_(torch.from_numpy(np.ascontiguousarray(_unlines(out, self.shape, 2))))

"""Multi-GPU ADI step: z-slab decomposition over the GPUs of one node.

There is no counterpart in the reference (single process, single thread).
One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch); rank r
owns planes ``[r*nz/W, (r+1)*nz/W)`` of the global ``(nz, ny, nx)`` grid, a
contiguous block in C order.

Per time step
  1. halo exchange: the top/bottom plane of ``T`` goes to the neighbouring
     ranks (``ny*nx*8`` bytes each way, send/recv pairs) - the explicit
     7-point stencil of stage 0 needs ``T`` at k+-1;
  2. x- and y-sweeps are slab-local (their lines do not cross slabs);
  3. z-sweep: every rank eliminates its own chunks of every z-line
     (``hs2_sweep_z_forward``), the 2 doubles per chunk per line that describe
     the chunk to its neighbours are all-gathered (``2*nz/chunk`` doubles per
     line in total - 0.4 % of the field for chunk = 32), and every rank
     back-substitutes with the rows of the global inverse interface operator
     (``hs2_sweep_z_backward``).  The field itself is never transposed or sent.

The result differs from the single-GPU path only by rounding (same chunked
algorithm; a single GPU applies the interface operator inside one kernel).

Two transports for the two exchanges:
  * peer memory (default on CUDA, ``PeerExchange``): every rank maps a mailbox of
    its neighbours (CUDA IPC over NVLink).  The z_forward kernel stores its
    interface values straight into the mailboxes of the slabs that need them,
    halo planes are copied there with cudaMemcpyAsync, and the GPUs order each
    other with flags in peer memory (stream-ordered signal / bounded wait
    kernels).  No collective call, no host synchronisation, one stream.
  * NCCL send/recv or all-gather (``HS2_DIST_P2P=0``, CPU/gloo tests, or when the
    interface band reaches more slabs than a kernel can address).
"""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi
from . import alternatingdirection_c_pyx as alternatingdirection
from . import crank_nicolson
from .plan import AdiPlan, interface_band


def slab_range(nz, rank, world):
    """Planes owned by ``rank``: equal slabs (nz must divide evenly)."""
    if nz % world != 0:
        raise ValueError("nz=%d is not divisible by the number of ranks (%d)" % (nz, world))
    h = nz // world
    return rank * h, (rank + 1) * h


class PeerExchange(object):
    """Mailbox of one rank in its own device memory, mapped by the ranks that
    write to it.  Layout (bytes):
      [0, 4096)   flags: Y-ready flag of source slab s at 8*s; halo-ready flags at
                  2048 (from the slab below) and 2056 (from above); status word at 3072
      halo_lo, halo_hi            one [ny][nx] plane each
      Y[parity]                   (2*hops+1) slab slots of 2*p_loc rows x n_lines doubles;
                                  slot i holds the rows of slab rank-hops+i
    Flags carry the step number (monotone), so nothing is ever reset."""

    HEADER = 4096
    OFF_FLAG_H_LO, OFF_FLAG_H_HI, OFF_STATUS = 2048, 2056, 3072

    def __init__(self, rank, world, hops, p_loc, ny, nx, group, device, fill_empty=False):
        self.rank, self.world, self.hops = rank, world, hops
        self.lib = _cabi.lib()
        self.peers = [r for d in range(1, hops + 1) for r in (rank - d, rank + d) if 0 <= r < world]
        if len(self.peers) > 6 or world > 240:
            raise NotImplementedError("interface band reaches %d slabs" % len(self.peers))
        self.n_lines = ny * nx
        self.plane_bytes = self.n_lines * 8
        self.slab_bytes = 2 * p_loc * self.n_lines * 8
        self.y_bytes = (2 * hops + 1) * self.slab_bytes
        self.off_halo_lo = self.HEADER
        self.off_halo_hi = self.HEADER + self.plane_bytes
        self.off_y = [self.HEADER + 2 * self.plane_bytes + par * self.y_bytes for par in (0, 1)]
        total = self.HEADER + 2 * self.plane_bytes + 2 * self.y_bytes
        # every step below is collective: a failure on any rank (no IPC support, no
        # peer access) makes ALL ranks give up, so that they fall back to NCCL together
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        self._own, self._opened, self.base = 0, [], {}
        with torch.cuda.device(device):
            ok = self.lib.hs2_peer_alloc(total, ctypes.byref(ptr), handle) == 0
        if ok:
            self._own = ptr.value
            self.base[rank] = ptr.value
            if fill_empty:
                # fused z sweep: an interface slot is "empty" until its value arrives (the data is its own flag)
                with torch.cuda.device(device):
                    _cabi.check(self.lib.hs2_peer_fill_empty(ptr.value + self.off_y[0], 2 * self.y_bytes // 8, None))
                    torch.cuda.synchronize()
        infos = [None] * world
        dist.all_gather_object(infos, (ok, handle.raw), group=group)
        ok = all(i[0] for i in infos)
        if ok:
            for r in self.peers:
                q = ctypes.c_void_p()
                with torch.cuda.device(device):
                    if self.lib.hs2_peer_open(ctypes.create_string_buffer(infos[r][1], 64), ctypes.byref(q)) != 0:
                        ok = False
                        break
                self.base[r] = q.value
                self._opened.append(q.value)
        oks = [None] * world
        dist.all_gather_object(oks, ok, group=group)
        if not all(oks):
            why = self.lib.hs2_last_error().decode("utf-8", "replace")
            self.close()
            raise NotImplementedError("peer memory unavailable on at least one rank (%s)" % why)
        dist.barrier(group=group)
        self.step_no = 0

    # virtual base address of the [2*P_glob][n_lines] interface array in rank r's mailbox
    def y_virtual(self, r, parity):
        return self.base[r] + self.off_y[parity] - (r - self.hops) * self.slab_bytes

    def y_rows_of(self, r, parity, src):
        """where slab ``src``'s rows live in rank r's mailbox"""
        return self.y_virtual(r, parity) + src * self.slab_bytes

    def status(self):
        out = torch.zeros(1, dtype=torch.int32)
        torch.cuda.synchronize()
        _cabi.check(self.lib.hs2_copy_async(out.data_ptr(), self._own + self.OFF_STATUS, 4, None))
        torch.cuda.synchronize()
        return int(out[0])

    def close(self):
        for q in self._opened:
            self.lib.hs2_peer_close(q)
        self._opened = []
        if self._own:
            self.lib.hs2_peer_free(self._own)
            self._own = 0


def _u64_list(vals):
    return (ctypes.c_uint64 * max(1, len(vals)))(*vals), len(vals)


class DistPlan(object):
    """Slab plan + the communication of one rank."""

    CHECK_EVERY = 64       # steps between reads of the peer-wait status word

    def __init__(self, plan, group=None):
        self.plan = plan
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.shape = plan.shape
        self.global_shape = (plan.slab["nz_global"],) + plan.shape[1:]
        self.k0 = plan.slab["k0"]
        self.chunk = plan.chunk[2][0]
        self.p_loc = plan.shape[0] // self.chunk
        self._band = None
        # pipelining pays once the exchange is large (measured: no gain at 2 ranks)
        self.pipeline = int(os.environ.get("HS2_DIST_PIPELINE", "1" if self.world <= 2 else "2"))
        self.min_lines = int(os.environ.get("HS2_DIST_MIN_LINES", "4096"))
        self.p2p_ranges = int(os.environ.get("HS2_DIST_P2P_RANGES", "2"))       # line ranges of the peer-memory z sweep (1 or 2)
        # one persistent kernel for the whole z sweep (forward, exchange through the mailboxes, backward);
        # HS2_DIST_Z_FUSED=0: forward / flag / backward launches per line range
        self.z_fused = os.environ.get("HS2_DIST_Z_FUSED", "0") != "0"
        self._bufs = {}
        self.use_p2p = os.environ.get("HS2_DIST_P2P", "1") != "0"
        # upper bound of a flag wait.  Ranks of one job drift apart by far more than a
        # step (I/O, compilation, a debugger) without anything being wrong, so the default
        # is long; a wait that does expire is fatal (check() raises, close() raises)
        self.p2p_timeout = float(os.environ.get("HS2_DIST_TIMEOUT_S", "1800"))
        self._px = None
        self.profile = None          # list: when set, _step_p2p appends 7 CUDA events per step

    @property
    def band(self):
        """half-width of the z interface operator in chunks: from the device plan once it exists (hs2_plan_build
        computed it), from the numpy statement of the tables otherwise (CPU runs)"""
        if self._band is None:
            if self.plan._handle is not None:
                self._band = int(self.plan.axis_info(2).band)
            else:
                self._band = interface_band(self.plan.chunk_tabs[2][1])
        return self._band

    @property
    def hops(self):
        """slabs whose interface values this slab needs: chunk p uses chunks p-1-band .. p+band"""
        return -(-(self.band + 1) // self.p_loc)

    # ------------------------------------------------------------- buffers
    def _buf(self, name, shape, like):
        b = self._bufs.get(name)
        if b is None or tuple(b.shape) != tuple(shape) or b.device != like.device:
            b = torch.empty(shape, dtype=torch.float64, device=like.device)
            self._bufs[name] = b
        return b

    def _peer(self, r):
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    # ------------------------------------------------- kernels (C ABI, CUDA)
    def _k_sweep_x(self, T_in, work, src, keep, halo_lo, halo_hi):
        p = self.plan
        _cabi.check(_cabi.lib().hs2_sweep_x(
            p._handle, T_in.data_ptr(), work.data_ptr(), ctypes.byref(src) if src is not None else None,
            halo_lo.data_ptr() if halo_lo is not None else None,
            halo_hi.data_ptr() if halo_hi is not None else None, self._stream(T_in)))

    def _k_sweep_y(self, work):
        _cabi.check(_cabi.lib().hs2_sweep_y(self.plan._handle, work.data_ptr(), self._stream(work)))

    def _k_z_forward(self, work, Y, line0, n_lines):
        _cabi.check(_cabi.lib().hs2_sweep_z_forward(self.plan._handle, work.data_ptr(), Y.data_ptr(), line0, n_lines,
                                                    self._stream(work)))

    def _k_z_backward(self, T_in, T_out, work, Yall, line0, n_lines):
        _cabi.check(_cabi.lib().hs2_sweep_z_backward(self.plan._handle, T_in.data_ptr(), T_out.data_ptr(),
                                                     work.data_ptr(), Yall.data_ptr(), line0, n_lines,
                                                     self._stream(work)))

    def _stream(self, t):
        return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)

    def _prepare(self, T_in):
        self.plan.ensure_device(T_in.device)

    # ----------------------------------------------- peer-memory transport
    def _peer_exchange(self, T_in):
        """The mailbox set-up is collective: the first step of every rank creates it."""
        if self._px is None and self.use_p2p and T_in.is_cuda and self.world > 1:
            try:
                ny, nx = self.shape[1:]
                self._px = PeerExchange(self.rank, self.world, self.hops, self.p_loc, ny, nx, self.group, T_in.device,
                                        fill_empty=self.z_fused)
            except NotImplementedError as exc:
                self.use_p2p = False
                self.p2p_unavailable = str(exc)
        return self._px

    def _step_p2p(self, px, T_in, T_out, work, src, keep):
        lib = _cabi.lib()
        st = self._stream(T_in)
        px.step_no += 1
        n = px.step_no
        par = n & 1
        me, lo, hi = self.rank, self.rank - 1, self.rank + 1
        own = px.base[me]
        sig, wait = [], []
        h = self.shape[0]
        ev = None
        if self.profile is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
            self.profile.append(ev)
            ev[0].record()
        # halo planes of T_in into the neighbours' mailboxes, then tell them
        if lo >= 0:
            _cabi.check(lib.hs2_copy_async(px.base[lo] + px.off_halo_hi, T_in.data_ptr(), px.plane_bytes, st))
            sig.append(px.base[lo] + px.OFF_FLAG_H_HI)
            wait.append(own + px.OFF_FLAG_H_LO)
        if hi < self.world:
            _cabi.check(lib.hs2_copy_async(px.base[hi] + px.off_halo_lo, T_in.data_ptr() + (h - 1) * px.plane_bytes,
                                           px.plane_bytes, st))
            sig.append(px.base[hi] + px.OFF_FLAG_H_LO)
            wait.append(own + px.OFF_FLAG_H_HI)
        arr, cnt = _u64_list(sig)
        _cabi.check(lib.hs2_flag_signal(arr, cnt, n, st))
        # interior planes need no halo: solve them while the neighbours' planes travel
        srcp = ctypes.byref(src) if src is not None else None
        h_lo = own + px.off_halo_lo if lo >= 0 else None
        h_hi = own + px.off_halo_hi if hi < self.world else None
        _cabi.check(lib.hs2_sweep_x_part(self.plan._handle, T_in.data_ptr(), work.data_ptr(), srcp, h_lo, h_hi, 1, st))
        if ev: ev[1].record()
        arr, cnt = _u64_list(wait)
        _cabi.check(lib.hs2_flag_wait(arr, cnt, n, self.p2p_timeout, own + px.OFF_STATUS, st))
        _cabi.check(lib.hs2_sweep_x_part(self.plan._handle, T_in.data_ptr(), work.data_ptr(), srcp, h_lo, h_hi, 2, st))
        if ev: ev[2].record()
        _cabi.check(lib.hs2_sweep_y(self.plan._handle, work.data_ptr(), st))
        if ev: ev[3].record()
        # z: eliminate the local chunks; their interface values go to every slab in reach.  Two ranges of lines:
        # the wait for range 0's rows is covered by the elimination of range 1, the wait for range 1's rows by
        # the back substitution of range 0 (flags of range i: slot 128 i + source slab)
        n_lines = self.shape[1] * self.shape[2]
        if self.z_fused:
            y_arr, cnt = _u64_list([px.y_rows_of(r, par, me) for r in px.peers])
            _cabi.check(lib.hs2_sweep_z_fused(self.plan._handle, T_in.data_ptr(), T_out.data_ptr(), work.data_ptr(),
                                              px.y_virtual(me, par), cnt, y_arr, self.p2p_timeout, own + px.OFF_STATUS, st))
            if ev:
                ev[4].record(); ev[5].record(); ev[6].record()
            return T_out
        ranges = self._line_ranges(n_lines, self.p2p_ranges if self.world <= 120 else 1)
        for i, (l0, nl) in enumerate(ranges):
            arr, cnt = _u64_list([px.y_rows_of(r, par, me) for r in px.peers])
            _cabi.check(lib.hs2_sweep_z_forward_push_cols(self.plan._handle, work.data_ptr(), px.y_rows_of(me, par, me), l0, nl,
                                                          cnt, arr, st))
            arr, cnt = _u64_list([px.base[r] + 8 * (128 * i + me) for r in px.peers])
            _cabi.check(lib.hs2_flag_signal(arr, cnt, n, st))
        if ev: ev[4].record()
        for i, (l0, nl) in enumerate(ranges):
            arr, cnt = _u64_list([own + 8 * (128 * i + r) for r in px.peers])
            _cabi.check(lib.hs2_flag_wait(arr, cnt, n, self.p2p_timeout, own + px.OFF_STATUS, st))
            if ev and i == 0: ev[5].record()
            _cabi.check(lib.hs2_sweep_z_backward_cols(self.plan._handle, T_in.data_ptr(), T_out.data_ptr(), work.data_ptr(),
                                                      px.y_virtual(me, par), l0, nl, st))
        if ev: ev[6].record()
        return T_out

    PHASES = ("halo_push+x_interior", "halo_wait+x_boundary", "y", "z_forward_push", "interface_wait(range 0)",
              "z_backward(+wait range 1)")

    def launches_per_step(self):
        """kernels of this library one step of this rank launches (interior rank, no source)"""
        n_lines = self.shape[1] * self.shape[2]
        if self._px is not None and self.z_fused:
            return 6       # flag signal (halo), x interior, flag wait, x boundary, y, fused z
        if self._px is not None:
            nr = len(self._line_ranges(n_lines, self.p2p_ranges if self.world <= 120 else 1))
            # flag signal (halo), x interior, flag wait, x boundary, y, then per range: forward, signal, wait, backward
            return 5 + 4 * nr
        return 2 + 2 * len(self._line_ranges(n_lines))

    def profile_ms(self):
        """mean milliseconds per phase over the profiled steps (synchronises)"""
        torch.cuda.synchronize()
        if not self.profile:
            return None
        n = len(self.profile)
        return {name: sum(e[i].elapsed_time(e[i + 1]) for e in self.profile) / n for i, name in enumerate(self.PHASES)}

    def check(self):
        """Raise if a peer-memory wait ever timed out (synchronises the device)."""
        if self._px is not None and self._px.status() != 0:
            raise RuntimeError("heatsim2_b200.dist: a peer did not deliver its data within %.0f s" % self.p2p_timeout)

    def close(self):
        if self._px is not None:
            torch.cuda.synchronize()
            bad = self._px.status() != 0
            self._px.close()
            self._px = None
            if bad:
                raise RuntimeError("heatsim2_b200.dist: a peer did not deliver its data within %.0f s; "
                                   "fields computed since are invalid" % self.p2p_timeout)

    # ------------------------------------------------------- communication
    def exchange_halos(self, T_in):
        """Send plane 0 down and plane -1 up; returns (halo_lo, halo_hi), None
        at the domain faces."""
        ny, nx = self.shape[1:]
        ops = []
        halo_lo = halo_hi = None
        if self.rank > 0:
            halo_lo = self._buf("halo_lo", (ny, nx), T_in)
            ops.append(dist.P2POp(dist.isend, T_in[0], self._peer(self.rank - 1), self.group))
            ops.append(dist.P2POp(dist.irecv, halo_lo, self._peer(self.rank - 1), self.group))
        if self.rank < self.world - 1:
            halo_hi = self._buf("halo_hi", (ny, nx), T_in)
            ops.append(dist.P2POp(dist.isend, T_in[-1], self._peer(self.rank + 1), self.group))
            ops.append(dist.P2POp(dist.irecv, halo_hi, self._peer(self.rank + 1), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return halo_lo, halo_hi

    def _line_ranges(self, n_lines, pipeline=None):
        """Split the z-lines into `pipeline` ranges (multiples of 16 lines)."""
        parts = max(1, min(self.pipeline if pipeline is None else pipeline, n_lines // self.min_lines))
        step = -(-n_lines // parts)
        step = -(-step // 16) * 16
        return [(l0, min(step, n_lines - l0)) for l0 in range(0, n_lines, step)]

    def _exchange_interface(self, Yall, own):
        """Deliver this slab's interface values to every slab that needs them
        and receive theirs into the matching slices of ``Yall``.  Chunk p needs
        the chunks p-1-band .. p+band, i.e. slabs within ``hops`` of its own;
        when that is (almost) everybody a single all-gather is used instead.
        Returns the outstanding requests."""
        if self.world == 1:
            return []
        if 2 * self.hops >= self.world - 1:
            return [dist.all_gather_into_tensor(Yall, own, group=self.group, async_op=True)]
        rows = 2 * self.p_loc
        ops = []
        for d in range(1, self.hops + 1):
            for r in (self.rank - d, self.rank + d):
                if 0 <= r < self.world:
                    ops.append(dist.P2POp(dist.isend, own, self._peer(r), self.group))
                    ops.append(dist.P2POp(dist.irecv, Yall[r * rows:(r + 1) * rows], self._peer(r), self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    # ----------------------------------------------------------------- step
    def step_device(self, T_in, T_out, t, dt, volumetric_elements, volumetric):
        """One time step of this rank's slab (tensors [nz/W, ny, nx] float64 on
        the plan's device; ``T_out`` may be ``T_in``).  Collective: every rank
        of the group must call it."""
        if tuple(T_in.shape) != self.shape:
            raise ValueError("slab has shape %r, got %r" % (self.shape, tuple(T_in.shape)))
        self._prepare(T_in)
        ny, nx = self.shape[1:]
        src = keep = None
        if volumetric is not None and len(volumetric):
            src, keep = self.plan._source(t, dt, volumetric_elements, volumetric)
        work = self._buf("work", self.shape, T_in)
        px = self._peer_exchange(T_in)
        if px is not None:
            self._step_p2p(px, T_in, T_out, work, src, keep)
            if keep is not None and len(keep) > 1:
                torch.cuda.current_stream(T_in.device).synchronize()
            return T_out
        halo_lo, halo_hi = self.exchange_halos(T_in)
        self._k_sweep_x(T_in, work, src, keep, halo_lo, halo_hi)
        self._k_sweep_y(work)
        # z-sweep, pipelined over ranges of lines: while the interface values of
        # one range travel over NVLink the next range is being eliminated
        n_lines = ny * nx
        ranges = self._line_ranges(n_lines)
        pending = []
        for (l0, nl) in ranges:
            Yall = self._buf("Yall%d" % l0, (self.world * 2 * self.p_loc, nl), T_in)
            own = Yall[self.rank * 2 * self.p_loc:(self.rank + 1) * 2 * self.p_loc]
            self._k_z_forward(work, own, l0, nl)
            pending.append((l0, nl, Yall, self._exchange_interface(Yall, own)))
        for (l0, nl, Yall, reqs) in pending:
            for req in reqs:
                req.wait()
            self._k_z_backward(T_in, T_out, work, Yall, l0, nl)
        if keep is not None and len(keep) > 1 and T_in.is_cuda:
            torch.cuda.current_stream(T_in.device).synchronize()
        return T_out

    def run_step(self, t, dt, Tarray, volumetric_elements, volumetric, out=None):
        if not isinstance(Tarray, torch.Tensor):
            raise TypeError("distributed run_adi_steps takes this rank's slab as a torch tensor")
        if Tarray.dtype != torch.float64:
            raise ValueError("Tarray must be float64")
        T_in = Tarray.contiguous()
        if out is not None:
            AdiPlan.check_out(out, T_in)
        T_out = torch.empty_like(T_in) if out is None else out
        if not T_in.is_cuda:
            return self.step_device(T_in, T_out, t, dt, volumetric_elements, volumetric)
        with torch.cuda.device(T_in.device):
            self.step_device(T_in, T_out, t, dt, volumetric_elements, volumetric)
            # a peer that never delivered must not go unnoticed: the waits give up after
            # p2p_timeout and set a status word; it is read (one device synchronisation)
            # every CHECK_EVERY steps, by check() and by close()
            if self._px is not None and self._px.step_no % self.CHECK_EVERY == 0:
                self.check()
        return T_out

    # bytes this rank sends per step (for NVLink accounting in bench.py)
    def comm_bytes_per_step(self):
        ny, nx = self.shape[1:]
        halo = ny * nx * 8 * ((self.rank > 0) + (self.rank < self.world - 1))
        if 2 * self.hops >= self.world - 1:
            peers = self.world - 1
        else:
            peers = sum(1 for d in range(1, self.hops + 1) for r in (self.rank - d, self.rank + d) if 0 <= r < self.world)
        if self._px is not None:
            peers = len(self._px.peers)
            mode = "peer-memory stores from the z_forward kernel to %d slab(s), flags in peer memory" % peers
        else:
            mode = "NCCL all-gather" if 2 * self.hops >= self.world - 1 else "NCCL send/recv, %d-hop neighbours" % self.hops
        return {"halo_send": halo, "interface_send": 2 * self.p_loc * ny * nx * 8 * peers, "interface_mode": mode}


def setup(z0, y0, x0, dz, dy, dx, nz, ny, nx, dt, materials, boundaries, volumetric,
          material_elements, boundary_z_elements, boundary_y_elements, boundary_x_elements, volumetric_elements,
          top_surface_y_curvatures=None, top_surface_x_curvatures=None, unaligned_anisotropic=False,
          group=None, device=None, plan_class=DistPlan):
    """Distributed counterpart of ``heatsim2_b200.setup``: same GLOBAL problem
    description on every rank; returns ``(ADI_params, ADI_steps)`` whose plan
    steps this rank's z-slab.  ``ADI_params.slab = (k0, k1)``.

    ``run_adi_steps(ADI_params, ADI_steps, t, dt, T_slab, vol_elements_slab,
    volumetric)`` then takes and returns the local slab."""
    nz, ny, nx = int(nz), int(ny), int(nx)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    k0, k1 = slab_range(nz, rank, world)
    class_id, coefs, volume_array, vol = crank_nicolson.compile_problem(
        z0, y0, x0, dz, dy, dx, nz, ny, nx, dt, materials, boundaries, volumetric, material_elements,
        boundary_z_elements, boundary_y_elements, boundary_x_elements, volumetric_elements,
        top_surface_y_curvatures, top_surface_x_curvatures, unaligned_anisotropic, device=device)
    slab_volume = volume_array[k0:k1] if np.ndim(volume_array) > 0 else volume_array    # per-cell volumes in curved mode
    plan = AdiPlan((k1 - k0, ny, nx), None, coefs, dt, slab_volume, volumetric_elements=vol[k0:k1],
                   materials=materials, slab=(k0, class_id))
    (ADI_params, ADI_steps) = alternatingdirection.adi_setup((nz, ny, nx), volume_array)
    ADI_params.plan = plan_class(plan, group)
    ADI_params.slab = (k0, k1)
    return (ADI_params, ADI_steps)

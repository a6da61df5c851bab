// kernels_xm.cu - stage 0 of the ADI step, z-marching variant of the folded
// x-sweep of kernels_xf.cu (same arithmetic, bit-identical results): the input
// field never passes through registers or L1 on its way in.  OPT-IN
// (HS2_FLAG_X_MARCH / HS2_X_KERNEL=march): measured on B200 it is 10 % slower
// than kernels_xf.cu at 512^3 (0.90 vs 0.82 ms) - it keeps one tile per SM where
// the folded kernel keeps four, and the number of solve chains in flight is what
// bounds the sweep (profiles/NOTES_r01.md, third part).
//
//   d1 = A^-1 [ 2 T + q ] - 2 T ,  q = M^-1 (Ly + Lz) T ,  A = I - 1/2 M^-1 Lx
//   (replaces B0.dot(T) + tridiagsolve of stage 0,
//    heatsim2/alternatingdirection_c_pyx.pyx:397-412)
//
// A persistent block owns work items (tile of R x-lines j0..j0+R-1, range of KR
// planes [ka, kb)) and marches through the planes of an item.  The rows
// j0-1..j0+R of every plane are fetched ONCE by the tensor copy engine
// (cp.async.bulk.tensor, completion on an mbarrier) into a ring of four
// shared-memory slices: while plane k is worked on, slices k-1, k, k+1 are
// resident (z neighbours and y halo come out of shared memory) and planes k+2,
// k+3 are in flight - two planes (2 x 40 KB at nx = 512, R = 8) of loads per SM
// are outstanding at any time without holding a single register.  The L2 -> SM
// traffic per cell falls from 37 B (kernels_xf.cu: z neighbours and the phase-3
// re-read come from L2 through L1) to 13-18 B, and phase 3 reads T from the slice.
//
//  phase 1  nx/2 column pairs x R/RPT row groups of threads (up to 512); 5-point
//           stencil from the three slices into registers, then - after a barrier,
//           because the solve buffer aliases the slice of plane k-1 - 2T + q to
//           the chunk-padded solve buffer (16-byte accesses);
//  phase 2  the first R*P threads: thread (line r, chunk p) runs the partitioned
//           solve in registers (chunk_core.cuh); factor tables and interface-
//           operator rows of the block's line class sit in shared memory (8-byte
//           broadcast loads), other classes read the global tables;
//  phase 3  d1 = w - 2T, T from the centre slice, coalesced 16-byte stores.
//
// R = 8 (default; one 512-thread block per SM at nx = 512) or 4 (HS2_XM_R=4: two
// 256-thread blocks per SM, tables from global memory).  The indexing and the
// copy/wait protocol are transcribed thread by thread in tests/xm_sim.py and
// checked on the CPU (tests/test_xmarch_sim.py).
//
// Not handled here (the caller falls back to kernels_xf.cu): steps with an
// active volumetric source, slab halos / plane parts of the multi-GPU path,
// nx > 512, fewer than R+2 rows, grids whose slices do not fit shared memory.
#include "x_common.cuh"
#include "tma_util.cuh"
#include <stdlib.h>

namespace {

constexpr int XM_SLOTS = 4;
constexpr int XM_PAD = 2;   // pad doubles per chunk in the solve buffer (as kernels_xf.cu, 8-line tiles)
constexpr int XM_TP = 2;    // pad doubles per chunk in the shared-memory factor tables

__host__ __device__ inline int xm_row_pitch(int P, int M) {
  int s = P * (M + XM_PAD);
  while ((s & 3) != 2) s += 2;
  return s;
}

struct XmGeom {
  int BX;           // columns per copy box (<= 256)
  int NXB;          // boxes per row
  int box_stride;   // doubles between the boxes of a slice (128-byte multiple)
  int slot_stride;  // doubles between slices
  int Sr;           // row pitch of the solve buffer
  int tab_in_smem;  // factor tables of the block's line class in shared memory
  int ge_in_smem;   // ... and its interface-operator rows
  int R;            // x-lines per tile (8 or 4)
  int buf_alias;    // solve buffer lives in the slice of plane k-1
  int RPT;          // rows per thread in phases 1 and 3 (2 or 4)
  int KR;           // planes per work item
  int tiles_y;
  int n_items;
};

// Shared-memory table accessor of this kernel: chunk_core.cuh's TabShared with a VOLATILE load.
// The block may keep no tables at all (tab_in_smem = 0); a plain asm is a pure function to the
// compiler, which then turns ld(select(c, tables, elsewhere)) into select(c, ld(tables), ld(elsewhere))
// and reads past the end of shared memory (found with compute-sanitizer; SASS: LDS.64 ahead of the SEL).
struct XmTabShared {
  uint32_t a;        // shared-window byte address of plane 0 at this thread's first row
  uint32_t pitch_b;  // bytes between planes
  __device__ __forceinline__ uint32_t plane(int pl) const { return a + pl * pitch_b; }
  __device__ __forceinline__ static double ld(uint32_t q, int t) {
    double x;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(q + 8u * (uint32_t)t));
    return x;
  }
};

__device__ __forceinline__ bool xm_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// a copy that never completes must not hang the GPU: give up after ~2 s
__device__ __forceinline__ void xm_wait(uint64_t *bar, uint32_t parity) {
  if (xm_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!xm_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

template <int M, int RPT, int R, typename CID>
__global__ void __launch_bounds__(512, 1)
sweep_xm_kernel(const __grid_constant__ CUtensorMap tmap, double *__restrict__ Wout, const CID *__restrict__ cid,
                const double *__restrict__ coef_g, int n_classes, int coef_in_smem,
                const uint32_t *__restrict__ line_id, const double *__restrict__ tab, int pitch,
                const double *__restrict__ GE, int nz, int ny, int nx, int P, int band, XmGeom g) {
  constexpr int RW = R + 2;
  constexpr int RG = R / RPT;           // row groups of phases 1 and 3
  extern __shared__ __align__(128) unsigned char xm_raw[];
  double *slots = reinterpret_cast<double *>(xm_raw);                          // [4][slot_stride]
  double *buf_own = slots + (size_t)XM_SLOTS * g.slot_stride;                  // [R][Sr] unless aliased
  double *Y = buf_own + (g.buf_alias ? 0 : R * g.Sr);                          // [2P][R]
  double *Es = Y + 2 * P * R;                                                  // [P][R]
  double *cfs = Es + P * R;                                                    // [n_classes][8]
  double *s_tab = cfs + (coef_in_smem ? n_classes * HS2_COEF_STRIDE : 0);      // [PLANES][P][M+TP]
  double *s_ge = s_tab + (g.tab_in_smem ? HS2_T_PLANES * P * (M + XM_TP) : 0); // [P][2P+2]
  uint64_t *bar = reinterpret_cast<uint64_t *>(s_ge + (g.ge_in_smem ? P * (2 * P + 2) : 0));
  HS2_MARK_DECL;
  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;
  const bool leader = tid == 0;
  const int64_t plane = (int64_t)ny * nx;
  const int BX = g.BX;

  if (leader) {
#pragma unroll
    for (int s = 0; s < XM_SLOTS; ++s) mbar_init(&bar[s], 1);
    fence_mbar_init();
  }
  if (coef_in_smem)
    for (int q = tid; q < n_classes * HS2_COEF_STRIDE; q += nthreads) cfs[q] = coef_g[q];
  const double *coef = coef_in_smem ? cfs : coef_g;
  // factor tables / interface operator of the class of this block's first line
  uint32_t lid_c = 0xffffffffu;
  if (g.tab_in_smem && (int)blockIdx.x < g.n_items) {
    const int kr0 = blockIdx.x / g.tiles_y, jt0 = blockIdx.x % g.tiles_y;
    lid_c = __ldg(line_id + (int64_t)(kr0 * g.KR) * ny + jt0 * R);
    const double *gt = tab + (int64_t)lid_c * HS2_T_PLANES * pitch;
    const int per_plane = P * (M + XM_TP);
    for (int e = tid; e < HS2_T_PLANES * per_plane; e += nthreads) {
      const int pl = e / per_plane, rem = e % per_plane;
      const int pp = rem / (M + XM_TP), t = rem % (M + XM_TP);
      const int row = pp * M + t;
      s_tab[e] = (t < M && row < pitch) ? gt[(int64_t)pl * pitch + row] : 0.0;
    }
    if (g.ge_in_smem) {
      const double *gg = GE + (int64_t)lid_c * P * 2 * P;
      for (int e = tid; e < P * (2 * P + 2); e += nthreads) {
        const int pp = e / (2 * P + 2), q = e % (2 * P + 2);
        s_ge[e] = q < 2 * P ? gg[pp * 2 * P + q] : 0.0;
      }
    }
  }
  __syncthreads();

  // phase-2 role (threads 0 .. R*P-1): chunk pc of line r2
  const bool solver = tid < R * P;
  const int r2 = tid % R;
  const int p2 = solver ? tid / R : P - 1;
  const int pc = p2;
  const int c0 = pc * M;
  const int rows = min(M, nx - c0);
  const bool full = rows == M;
  const int mine_off = r2 * g.Sr + pc * (M + XM_PAD);
  XmTabShared ts;
  ts.a = g.tab_in_smem ? smem_u32(s_tab + pc * (M + XM_TP)) : smem_u32(slots);
  ts.pitch_b = g.tab_in_smem ? (uint32_t)(P * (M + XM_TP)) * 8u : 0u;

  // phase-1/3 role (threads 0 .. RG*nx/2-1): column pair ic, ic+1 of rows r0 .. r0+RPT-1
  const int cpt = nx >> 1;
  const int rg = tid / cpt;
  const bool stencil = rg < RG;
  const int ic = stencil ? 2 * (tid - rg * cpt) : 0;
  const int r0 = stencil ? rg * RPT : 0;
  const int so = (ic / BX) * g.box_stride + (ic % BX);
  const int bo = ic + XM_PAD * (ic / M);

  const uint32_t slice_bytes = (uint32_t)g.NXB * RW * BX * sizeof(double);
  uint32_t phase_bits = 0;   // bit s: parity the next wait on slot s uses

  for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
    const int kr = item / g.tiles_y;
    const int j0 = (item % g.tiles_y) * R;
    const int ka = kr * g.KR;
    const int kb = min(nz, ka + g.KR);
    const int n_it = kb - ka;
    const int nrows = min(R, ny - j0);
    // load sequence q = 0 .. n_it+1  <->  plane ka-1+q, slot q & 3
    auto needed = [&](int q) {
      const int pl = ka - 1 + q;
      return pl >= 0 && pl < nz && q <= n_it + 1;
    };
    auto issue = [&](int q) {   // leader only
      const int s = q & (XM_SLOTS - 1);
      mbar_expect_tx(&bar[s], slice_bytes);
      double *dst = slots + (size_t)s * g.slot_stride;
      for (int xb = 0; xb < g.NXB; ++xb)
        tma_load_3d(dst + (size_t)xb * g.box_stride, &tmap, &bar[s], xb * BX, j0 - 1, ka - 1 + q);
    };
    auto wait_slot = [&](int q) {
      const int s = q & (XM_SLOTS - 1);
      xm_wait(&bar[s], (phase_bits >> s) & 1u);
      phase_bits ^= 1u << s;
    };
    if (leader) {
#pragma unroll
      for (int q = 0; q < 3; ++q)
        if (needed(q)) issue(q);
    }
    // window rows r0-1 .. r0+RPT of the centre slice, clamped to the plane like kernels_xf.cu
    int rq[RPT + 2];
#pragma unroll
    for (int q = 0; q < RPT + 2; ++q) {
      int j = j0 + r0 + q - 1;
      j = j < 0 ? 0 : (j > ny - 1 ? ny - 1 : j);
      rq[q] = (j - (j0 - 1)) * BX;
    }
    // class ids of plane ka (this thread's column pair, rows clamped)
    uint32_t idc[RPT], idn[RPT];
    {
      const CID *cidk = cid + (int64_t)ka * plane;
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int j = min(j0 + r0 + r, ny - 1);
        const CID *a = cidk + ((int64_t)j * nx + ic);
        if (sizeof(CID) == 1)
          idc[r] = *reinterpret_cast<const uint16_t *>(a);
        else
          idc[r] = *reinterpret_cast<const uint32_t *>(a);
        idn[r] = 0;
      }
    }

    for (int it = 0; it < n_it; ++it) {
      const int k = ka + it;
      const int64_t kbase = (int64_t)k * plane;
      if (leader && needed(it + 3)) issue(it + 3);
      const uint32_t lid = __ldg(line_id + (int64_t)k * ny + j0 + (r2 < nrows ? r2 : 0));
      if (it + 1 < n_it) {   // class ids of the next plane: in flight during this one
        const CID *cidk = cid + kbase + plane;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          const int j = min(j0 + r0 + r, ny - 1);
          const CID *a = cidk + ((int64_t)j * nx + ic);
          if (sizeof(CID) == 1)
            idn[r] = *reinterpret_cast<const uint16_t *>(a);
          else
            idn[r] = *reinterpret_cast<const uint32_t *>(a);
        }
      }
      HS2_MARK(8);
      if (it == 0) {
        if (needed(0)) wait_slot(0);
        wait_slot(1);
      }
      if (k + 1 < nz) wait_slot(it + 2);
      HS2_MARK(0);
      const double *Cn = slots + (size_t)((it + 1) & (XM_SLOTS - 1)) * g.slot_stride;
      const double *Lo = k > 0 ? slots + (size_t)(it & (XM_SLOTS - 1)) * g.slot_stride : Cn;
      const double *Hi = k + 1 < nz ? slots + (size_t)((it + 2) & (XM_SLOTS - 1)) * g.slot_stride : Cn;
      // solve buffer: its own region, or (buf_alias) the slice of plane k-1, which
      // nobody reads after phase 1 and no copy targets before the next iteration
      double *buf = g.buf_alias ? slots + (size_t)(it & (XM_SLOTS - 1)) * g.slot_stride : buf_own;

      // ------------------------------------------------ phase 1: 2T + q
      double2 outv[RPT];
      if (stencil) {
        const double *cc = Cn + so;
        const double *zl = Lo + so;
        const double *zh = Hi + so;
        double2 tc[RPT + 2];
#pragma unroll
        for (int q = 0; q < RPT + 2; ++q) tc[q] = *reinterpret_cast<const double2 *>(cc + rq[q]);
        int last_id = -1;
        double2 cy = make_double2(0, 0), cz = cy;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          const double2 zm = *reinterpret_cast<const double2 *>(zl + (r0 + r + 1) * BX);
          const double2 zp = *reinterpret_cast<const double2 *>(zh + (r0 + r + 1) * BX);
          const int q = r + 1;
          const double2 t0 = tc[q];
          const int id0 = sizeof(CID) == 1 ? (int)(idc[r] & 0xff) : (int)(idc[r] & 0xffff);
          const int id1 = sizeof(CID) == 1 ? (int)((idc[r] >> 8) & 0xff) : (int)((idc[r] >> 16) & 0xffff);
          double out[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int idc_e = e ? id1 : id0;
            if (idc_e != last_id) {
              const double2 *c2 = reinterpret_cast<const double2 *>(coef + idc_e * HS2_COEF_STRIDE);
              cy = c2[1];
              cz = c2[2];
              last_id = idc_e;
            }
            const double tt = e ? t0.y : t0.x;
            const double vym = e ? tc[q - 1].y : tc[q - 1].x;
            const double vyp = e ? tc[q + 1].y : tc[q + 1].x;
            const double vzm = e ? zm.y : zm.x;
            const double vzp = e ? zp.y : zp.x;
            double rr = cy.x * (vym - tt);
            rr = fma(cy.y, vyp - tt, rr);
            rr = fma(cz.x, vzm - tt, rr);
            rr = fma(cz.y, vzp - tt, rr);
            out[e] = fma(2.0, tt, rr);
          }
          outv[r] = make_double2(out[0], out[1]);
        }
      }
      if (g.buf_alias) __syncthreads();   // every read of slice k-1 is done before the solve buffer overwrites it
      if (stencil) {
        double *bcol = buf + bo;
#pragma unroll
        for (int r = 0; r < RPT; ++r) *reinterpret_cast<double2 *>(bcol + (r0 + r) * g.Sr) = outv[r];
      }
      HS2_MARK(1);
      __syncthreads();

      // ------------------------------------------------ phase 2: solve along x
      double *mine = buf + mine_off;
      const bool tab_s = g.tab_in_smem && lid == lid_c;
      TabGlobal tg;
      tg.b = tab + ((int64_t)lid * HS2_T_PLANES) * pitch + c0;
      tg.pitch = pitch;
      const double *ge = (tab_s && g.ge_in_smem) ? s_ge + pc * (2 * P + 2) : GE + ((int64_t)lid * P + pc) * (2 * P);
      double v[M];
      if (solver) {
#pragma unroll
        for (int t = 0; t < M; t += 2) {
          double2 x = make_double2(0.0, 0.0);
          if (full || t < rows) x = *reinterpret_cast<const double2 *>(mine + t);
          v[t] = x.x;
          v[t + 1] = x.y;
        }
        double yf, last;
        if (tab_s) {
          if (full)
            yf = chunk_fwd<M, true>(v, ts, M, &last);
          else
            yf = chunk_fwd<M, false>(v, ts, rows, &last);
        } else {
          if (full)
            yf = chunk_fwd<M, true>(v, tg, M, &last);
          else
            yf = chunk_fwd<M, false>(v, tg, rows, &last);
        }
        Y[(2 * p2) * R + r2] = yf;
        Y[(2 * p2 + 1) * R + r2] = last;
      }
      HS2_MARK(2);
      __syncthreads();
      double E = 0.0;
      if (solver) {
        E = chunk_interface(ge, Y, P, R, r2, pc, band);
        Es[p2 * R + r2] = E;
      }
      HS2_MARK(3);
      __syncthreads();
      if (solver) {
        const double alpha = p2 > 0 ? Es[(p2 - 1) * R + r2] : 0.0;
        if (tab_s) {
          if (full)
            chunk_bwd<M, true>(v, ts, M, alpha, E);
          else
            chunk_bwd<M, false>(v, ts, rows, alpha, E);
        } else {
          if (full)
            chunk_bwd<M, true>(v, tg, M, alpha, E);
          else
            chunk_bwd<M, false>(v, tg, rows, alpha, E);
        }
#pragma unroll
        for (int t = 0; t < M; t += 2)
          if (full || t < rows) *reinterpret_cast<double2 *>(mine + t) = make_double2(v[t], v[t + 1]);
      }
      HS2_MARK(4);
      __syncthreads();

      // ------------------------------------------------ phase 3: d1 = w - 2T
      if (stencil) {
        const double *cc = Cn + so;
        const double *bcol = buf + bo;
        double *Wk = Wout + kbase + (int64_t)j0 * nx + ic;
        double2 t0[RPT], w[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          t0[r] = *reinterpret_cast<const double2 *>(cc + (r0 + r + 1) * BX);
          w[r] = *reinterpret_cast<const double2 *>(bcol + (r0 + r) * g.Sr);
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r)
          if (r0 + r < nrows)
            *reinterpret_cast<double2 *>(Wk + (int64_t)(r0 + r) * nx) =
                make_double2(fma(-2.0, t0[r].x, w[r].x), fma(-2.0, t0[r].y, w[r].y));
      }
#pragma unroll
      for (int r = 0; r < RPT; ++r) idc[r] = idn[r];
      // the buffer's slice is the target of the next iteration's copy: order this
      // thread's shared-memory accesses before the copy engine's writes
      if (g.buf_alias) fence_proxy_async();
      HS2_MARK(5);
      __syncthreads();   // buf, Y, Es and the oldest slice are rewritten by the next plane
    }
  }
}

// shared-memory layout and block shape for a plan; false: outside what the kernel handles
template <int M>
bool xm_geometry(const hs2_plan *p, XmGeom *g, size_t *smem_out, int *threads_out, int *coef_in_smem_out) {
  const hs2_plan_desc &d = p->d;
  const hs2_axis_tables &ax = d.axis[0];
  // HS2_XM_R: x-lines per tile, 8 (one 512-thread block per SM at nx = 512) or 4 (two 256-thread
  // blocks per SM, factor tables from global memory / L1); HS2_XM_TAB=0/1 overrides where the tables live
  const char *r_env = getenv("HS2_XM_R");
  const int R = (r_env && atoi(r_env) == 4) ? 4 : 8;
  const int RW = R + 2;
  const char *t_env = getenv("HS2_XM_TAB");
  const bool want_tab = t_env ? atoi(t_env) != 0 : R == 8;
  g->R = R;
  const int P = ax.n_chunks;
  const int nx = (int)d.nx, ny = (int)d.ny;
  if (R * P > 256 || (nx & 1) || ny < RW) return false;
  if (!ax.d_tab || !ax.d_GE || ax.pitch <= 0) return false;
  // phases 1 and 3: nx/2 column pairs x RG row groups of RPT rows; as many threads as fit a block
  const int cpt = nx / 2;
  if (2 * cpt > 512) return false;
  g->RPT = 4 * cpt <= 512 ? 2 : 4;
  int threads = (R / g->RPT) * cpt;
  if (threads < R * P) threads = R * P;
  threads = (threads + 31) / 32 * 32;
  if (threads > 512) return false;
  g->BX = nx <= 256 ? nx : 256;
  g->NXB = (nx + g->BX - 1) / g->BX;
  g->box_stride = ((RW * g->BX + 15) / 16) * 16;
  g->slot_stride = g->NXB * g->box_stride;
  g->Sr = xm_row_pitch(P, M);
  g->buf_alias = R * g->Sr <= g->slot_stride ? 1 : 0;
  const int coef_in_smem = d.n_classes <= 256 ? 1 : 0;
  const size_t base = ((size_t)XM_SLOTS * g->slot_stride + (g->buf_alias ? 0 : (size_t)R * g->Sr) + 3 * (size_t)P * R +
                       (coef_in_smem ? (size_t)d.n_classes * HS2_COEF_STRIDE : 0)) * sizeof(double) +
                      XM_SLOTS * sizeof(uint64_t);
  const size_t tabs = (size_t)HS2_T_PLANES * P * (M + XM_TP) * sizeof(double);
  const size_t ges = (size_t)P * (2 * P + 2) * sizeof(double);
  g->tab_in_smem = (want_tab && base + tabs <= (size_t)p->max_smem_optin) ? 1 : 0;
  g->ge_in_smem = (g->tab_in_smem && base + tabs + ges <= (size_t)p->max_smem_optin) ? 1 : 0;
  const size_t smem = base + (g->tab_in_smem ? tabs : 0) + (g->ge_in_smem ? ges : 0);
  if (smem > (size_t)p->max_smem_optin) return false;
  *smem_out = smem;
  *threads_out = threads;
  *coef_in_smem_out = coef_in_smem;
  return true;
}

template <int M, int RPT, int R, typename CID>
int launch_xm_k(hs2_plan *p, const XmGeom &g0, size_t smem, int threads, int coef_in_smem, const double *T, double *W,
                cudaStream_t st, bool *done) {
  const hs2_plan_desc &d = p->d;
  const hs2_axis_tables &ax = d.axis[0];
  constexpr int RW = R + 2;
  const int P = ax.n_chunks;
  const int nx = (int)d.nx, ny = (int)d.ny, nz = (int)d.nz;
  XmGeom g = g0;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (!hs2_encode_tmap_f64_3d(&tmap, T, (uint64_t)nx, (uint64_t)ny, (uint64_t)nz, (uint32_t)g.BX, RW, 1)) return HS2_OK;
  auto kern = sweep_xm_kernel<M, RPT, R, CID>;
  HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  int occ = 0;
  HS2_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  if (occ < 1) return HS2_OK;
  const char *kr_env = getenv("HS2_XM_KR");   // planes per work item (read per launch: tests vary it)
  int KR = kr_env ? atoi(kr_env) : 32;
  if (KR < 1) KR = 1;
  if (KR > nz) KR = nz;
  g.KR = KR;
  g.tiles_y = (ny + R - 1) / R;
  const int64_t n_items = (int64_t)((nz + KR - 1) / KR) * g.tiles_y;
  if (n_items >= ((int64_t)1 << 31)) return HS2_OK;
  g.n_items = (int)n_items;
  int64_t blocks = (int64_t)p->sm_count * occ;
  if (blocks > n_items) blocks = n_items;
  kern<<<(unsigned)blocks, threads, smem, st>>>(tmap, W, (const CID *)d.d_class_id, d.d_class_coef, d.n_classes,
                                                 coef_in_smem, ax.d_line_id, ax.d_tab, ax.pitch, ax.d_GE, nz, ny, nx, P,
                                                 ax.band, g);
  HS2_CUDA_CHECK(cudaGetLastError());
  *done = true;
  return HS2_OK;
}

template <int M, typename CID>
int launch_xm(hs2_plan *p, const double *T, double *W, cudaStream_t st, bool *done) {
  XmGeom g;
  size_t smem = 0;
  int threads = 0, coef_in_smem = 0;
  if (!xm_geometry<M>(p, &g, &smem, &threads, &coef_in_smem)) return HS2_OK;
  if (g.R == 4) {
    if (g.RPT == 2) return launch_xm_k<M, 2, 4, CID>(p, g, smem, threads, coef_in_smem, T, W, st, done);
    return launch_xm_k<M, 4, 4, CID>(p, g, smem, threads, coef_in_smem, T, W, st, done);
  }
  if (g.RPT == 2) return launch_xm_k<M, 2, 8, CID>(p, g, smem, threads, coef_in_smem, T, W, st, done);
  return launch_xm_k<M, 4, 8, CID>(p, g, smem, threads, coef_in_smem, T, W, st, done);
}

template <typename CID>
int dispatch_xm(hs2_plan *p, const double *T, double *W, cudaStream_t st, bool *done) {
  switch (p->d.axis[0].chunk) {
    case 8: return launch_xm<8, CID>(p, T, W, st, done);
    case 16: return launch_xm<16, CID>(p, T, W, st, done);
    case 32: return launch_xm<32, CID>(p, T, W, st, done);
  }
  return HS2_OK;
}

}  // namespace

// plan-level applicability (no launch): flag set, tile kernels usable, slices fit
bool hs2_tile_xm_supported(const hs2_plan *p) {
  if (!(p->d.flags & HS2_FLAG_X_MARCH) || !hs2_tile_xf_supported(p)) return false;
  XmGeom g;
  size_t smem;
  int threads, cs;
  switch (p->d.axis[0].chunk) {
    case 8: return xm_geometry<8>(p, &g, &smem, &threads, &cs);
    case 16: return xm_geometry<16>(p, &g, &smem, &threads, &cs);
    case 32: return xm_geometry<32>(p, &g, &smem, &threads, &cs);
  }
  return false;
}

// Tries the z-marching kernel; *done stays false when the plan/grid is outside
// what it handles (the caller then runs kernels_xf.cu).
int hs2_tile_sweep_xm(hs2_plan *p, const double *T, double *W, cudaStream_t st, bool *done) {
  *done = false;
  if (p->d.class_id_bytes == 1) return dispatch_xm<uint8_t>(p, T, W, st, done);
  return dispatch_xm<uint16_t>(p, T, W, st, done);
}

#ifdef HS2_PHASE_TIMING
extern "C" int hs2_debug_phase_xm(unsigned long long *out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_hs2_phase, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_hs2_phase, z, sizeof(z));
  }
  return 0;
}
#endif

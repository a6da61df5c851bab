// kernels_xf.cu - stage 0 of the ADI step in one kernel, "folded" form: the
// explicit x-term of the right hand side is absorbed into the implicit solve.
//
//   stage 0:  A d1 = M^-1 [ (Lx + Ly + Lz) T + D s ] ,   A = I - 1/2 M^-1 Lx
//   M^-1 Lx T = 2 T - 2 A T      =>      d1 = A^-1 [ 2 T + q ] - 2 T ,
//   q = M^-1 [ (Ly + Lz) T + D s ]
//
// so phase 1 needs a 5-point stencil only (no x neighbours, no unaligned
// loads) and the tile is the same one the solve works on.  This is the
// arithmetic of the reference's own stage 0, which also solves for the full
// field and not for the increment (heatsim2/alternatingdirection_c_pyx.pyx:
// 397-412: B0.dot(T) + Dvec*src, tridiagsolve); the increment is formed with one
// fused multiply-add per cell when the result is stored.  Rounding error is
// eps * cond(A) * |T| per step, like the reference's.
//
// Replaces B0.dot(T) + Dvec*src + tridiagsolve of stage 0: neither the CSR
// matrix nor the right hand side ever exist in HBM.  HBM traffic per cell: read
// T once (y/z neighbours and the phase-3 re-read are L1/L2 hits), write d1
// once, + 1-2 B class id.
//
// A block owns R = 8 consecutive x-lines (rows j0..j0+7 of plane k) per tile
// and walks over tiles (persistent grid).
//  phase 1  threads own column pairs (16-byte accesses, coalesced along x), keep
//           a sliding y-window in registers, fetch the z neighbours from L2 and
//           write 2T + q to shared memory (chunk-padded layout, 16-byte stores);
//  phase 2  thread (r, p) takes chunk p (M consecutive cells) of line r into
//           registers with 16-byte shared loads and runs the partitioned solve
//           (chunk_core.cuh algebra).  The factor tables are read from a
//           chunk-interleaved copy [plane][row pair][chunk] so that the four
//           chunks a warp works on share one 64-byte segment per load;
//  phase 3  the solution goes back through shared memory, d1 = w - 2T, coalesced
//           16-byte stores.
// Shared-memory layout: cell (r, i) at r*Sr + i + 2*(i/M) doubles; two pad words
// per chunk keep every access 16-byte aligned and Sr/2 odd makes the
// chunk-major accesses of phase 2 (a quarter warp = 8 lines of one chunk)
// conflict free.
#include "x_common.cuh"
#include "tma_util.cuh"
#include <stdlib.h>

namespace {

// pad words per chunk in the shared-memory tile: 2 for 8-line tiles; 4-line
// tiles put two chunks into a quarter warp, which needs the chunk stride to be
// 4 (mod 8) 16-byte units
__host__ __device__ constexpr int xf_pad(int R, int M) { return R >= 8 ? 2 : (M == 8 ? 0 : 8); }

__host__ __device__ inline int xf_row_pitch(int P, int M, int pad) {
  int s = P * (M + pad);
  while ((s & 3) != 2) s += 2;
  return s;
}

// chunk-interleaved factor tables of one unique line: double2 element
// (plane, row pair t2, chunk p) at ((plane*(M/2) + t2)*P + p)
template <int M>
struct XTab {
  const double2 *b;  // already offset to this thread's chunk
  int P;
  __device__ __forceinline__ double2 pair(int plane, int t2) const { return __ldg(b + (plane * (M / 2) + t2) * P); }
  __device__ __forceinline__ double one(int plane, int t) const {
    const double2 c = pair(plane, t >> 1);
    return (t & 1) ? c.y : c.x;
  }
};

template <int M>
__device__ __forceinline__ double xf_forward_full(double (&v)[M], const XTab<M> &tb) {
#pragma unroll
  for (int t = 0; t < M; t += 2) {
    const double2 c = tb.pair(HS2_T_INV, t / 2);
    v[t] *= c.x;
    v[t + 1] *= c.y;
  }
  double prev = 0.0;
#pragma unroll
  for (int t = 0; t < M; t += 2) {
    const double2 c = tb.pair(HS2_T_F, t / 2);
    prev = fma(-c.x, prev, v[t]);
    v[t] = prev;
    prev = fma(-c.y, prev, v[t + 1]);
    v[t + 1] = prev;
  }
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int t = 0; t < M; t += 2) {
    const double2 c = tb.pair(HS2_T_C, t / 2);
    a0 = fma(c.x, v[t], a0);
    a1 = fma(c.y, v[t + 1], a1);
  }
  return a0 + a1;
}

template <int M>
__device__ __forceinline__ double xf_forward_short(double (&v)[M], const XTab<M> &tb, int rows, double *last) {
  double prev = 0.0, yf = 0.0;
#pragma unroll
  for (int t = 0; t < M; ++t) {
    if (t < rows) {
      prev = fma(-tb.one(HS2_T_F, t), prev, v[t] * tb.one(HS2_T_INV, t));
      v[t] = prev;
      yf = fma(tb.one(HS2_T_C, t), prev, yf);
    }
  }
  *last = prev;
  return yf;
}

template <int M>
__device__ __forceinline__ void xf_backward_full(double (&v)[M], const XTab<M> &tb, double alpha, double E) {
#pragma unroll
  for (int t = 0; t < M; t += 2) {
    const double2 c = tb.pair(HS2_T_S, t / 2);
    v[t] = fma(-alpha, c.x, v[t]);
    v[t + 1] = fma(-alpha, c.y, v[t + 1]);
  }
  double nxt = E;
  v[M - 1] = E;
#pragma unroll
  for (int t = M - 2; t >= 0; t -= 2) {
    const double2 c = tb.pair(HS2_T_CP, t / 2);  // (cp[t], cp[t+1])
    if (t + 1 < M - 1) {
      nxt = fma(-c.y, nxt, v[t + 1]);
      v[t + 1] = nxt;
    }
    nxt = fma(-c.x, nxt, v[t]);
    v[t] = nxt;
  }
}

template <int M>
__device__ __forceinline__ void xf_backward_short(double (&v)[M], const XTab<M> &tb, int rows, double alpha, double E) {
  double nxt = E;
#pragma unroll
  for (int t = M - 1; t >= 0; --t) {
    if (t < rows) {
      if (t < rows - 1) nxt = fma(-tb.one(HS2_T_CP, t), nxt, fma(-alpha, tb.one(HS2_T_S, t), v[t]));
      v[t] = nxt;
    }
  }
}

template <int M, int R, typename CID>
__global__ void __launch_bounds__(256, 2)
sweep_xf_kernel(const double *__restrict__ T, double *__restrict__ Wout, const CID *__restrict__ cid,
                const double *__restrict__ coef_g, int n_classes, int coef_in_smem, const uint8_t *__restrict__ vol,
                SrcTab st, const double *__restrict__ dense, const double *__restrict__ halo_lo,
                const double *__restrict__ halo_hi, const uint32_t *__restrict__ line_id,
                const double2 *__restrict__ xtab, const double *__restrict__ GE, int nz, int ny, int nx, int P, int band,
                int tiles_y, int n_tiles, int pf, int part) {
  extern __shared__ __align__(16) double sm[];
  constexpr int PAD = xf_pad(R, M);
  const int Sr = xf_row_pitch(P, M, PAD);
  double *buf = sm;            // [R][Sr]
  double *Y = buf + R * Sr;    // [2P][R]
  double *Es = Y + 2 * P * R;  // [P][R]
  double *cfs = Es + P * R;    // [n_classes][8] when coef_in_smem
  HS2_MARK_DECL;
  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;
  const int64_t plane = (int64_t)ny * nx;
  if (coef_in_smem)
    for (int q = tid; q < n_classes * HS2_COEF_STRIDE; q += nthreads) cfs[q] = coef_g[q];
  __syncthreads();
  const double *coef = coef_in_smem ? cfs : coef_g;
  const bool has_src = dense != nullptr || st.n > 0;
  // phase-2 role of this thread: chunk pc of line r
  const int r2 = tid % R;
  const int p2 = tid / R;
  const int pc = p2 < P ? p2 : P - 1;
  const int c0 = pc * M;
  const int rows = min(M, nx - c0);
  const bool full = rows == M;
  double *mine = buf + r2 * Sr + pc * (M + PAD);

  // part 0: every plane; 1: interior planes 1..nz-2 (need no halo); 2: the two
  // boundary planes 0 and nz-1 (a slab's first/last plane read the neighbours' halos)
  const int n_work = part == 0 ? n_tiles : (part == 1 ? n_tiles - 2 * tiles_y : 2 * tiles_y);
  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
    const int tile = part == 0 ? work : (part == 1 ? work + tiles_y : (work < tiles_y ? work : n_tiles - 2 * tiles_y + work));
    const int k = tile / tiles_y;
    const int j0 = (tile % tiles_y) * R;
    const int64_t kbase = (int64_t)k * plane;
    const int nrows = min(R, ny - j0);
    // unique-line id of this thread's phase-2 line: issued now, needed after phase 1
    const uint32_t lid = __ldg(line_id + (int64_t)k * ny + j0 + (r2 < nrows ? r2 : 0));
    if (pf) {
      // the only rows of a tile nobody has touched yet are its z+ neighbours:
      // ask L2 for those of this block's NEXT tile now (128-byte lines)
      const int tn = tile + gridDim.x;
      if (part == 0 && tn < n_tiles) {
        const int kn = tn / tiles_y, jn = (tn % tiles_y) * R;
        if (kn + 1 < nz) {
          const double *nxt = T + (int64_t)(kn + 1) * plane + (int64_t)jn * nx;
          const int n_el = min(R, ny - jn) * nx;
          for (int e = tid * 16; e < n_el; e += nthreads * 16) prefetch_l2(nxt + e);
        }
      }
    }

    // ------------------------------------------------ phase 1: 2T + q
    // z neighbours: the planes below/above, the neighbouring slab's halo plane,
    // or (domain face, conductance 0) the cell itself
    const double *zlo = k > 0 ? T + kbase - plane : (halo_lo ? halo_lo : T + kbase);
    const double *zhi = k < nz - 1 ? T + kbase + plane : (halo_hi ? halo_hi : T + kbase);
    const double *Tk = T + kbase;
    const CID *cidk = cid + kbase;
    constexpr int NB = R >= 8 ? 2 : 1;  // register batches per column pair
    constexpr int RH = R / NB;          // rows per batch
    // in-plane offsets (32-bit: ny*nx < 2^31) of the window rows j0-1 .. j0+R, clamped to the plane
    int ro[R + 2];
#pragma unroll
    for (int q = 0; q < R + 2; ++q) {
      int j = j0 + q - 1;
      j = j < 0 ? 0 : (j > ny - 1 ? ny - 1 : j);
      ro[q] = j * nx;
    }
    for (int i = 2 * tid; i < nx; i += 2 * nthreads) {
      double *bcol = buf + i + PAD * (i / M);
      int last_id = -1;
      double2 cy = make_double2(0, 0), cz = cy;
      double csrc = 0.0;
      double2 tc[R + 2];  // y-window: tc[q] = row j0 + q - 1
#pragma unroll
      for (int h = 0; h < NB; ++h) {
        const int rb = h * RH;
        double2 zm[RH], zp[RH];
        int id[RH];
        // issue every load of the batch before any arithmetic; rows already
        // in the window (loaded by batch 0) are not fetched again
#pragma unroll
        for (int q = (h == 0 ? 0 : RH + 2); q <= rb + RH + 1; ++q)
          tc[q] = *reinterpret_cast<const double2 *>(Tk + (ro[q] + i));
#pragma unroll
        for (int r = 0; r < RH; ++r) {
          const int o = ro[rb + r + 1] + i;
          zm[r] = *reinterpret_cast<const double2 *>(zlo + o);
          zp[r] = *reinterpret_cast<const double2 *>(zhi + o);
          if (sizeof(CID) == 1)
            id[r] = *reinterpret_cast<const uint16_t *>(cidk + o);
          else
            id[r] = (int)*reinterpret_cast<const uint32_t *>(cidk + o);
        }
#pragma unroll
        for (int r = 0; r < RH; ++r) {
          const int q = rb + r + 1;
          const double2 t0 = tc[q];
          const int id0 = sizeof(CID) == 1 ? (id[r] & 0xff) : (id[r] & 0xffff);
          const int id1 = sizeof(CID) == 1 ? ((id[r] >> 8) & 0xff) : ((id[r] >> 16) & 0xffff);
          double out[2];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int idc = c ? id1 : id0;
            if (idc != last_id) {  // interior cells share one class: usually not taken
              const double2 *c2 = reinterpret_cast<const double2 *>(coef + idc * HS2_COEF_STRIDE);
              cy = c2[1];
              cz = c2[2];
              csrc = c2[3].x;
              last_id = idc;
            }
            const double tt = c ? t0.y : t0.x;
            const double vym = c ? tc[q - 1].y : tc[q - 1].x;
            const double vyp = c ? tc[q + 1].y : tc[q + 1].x;
            const double vzm = c ? zm[r].y : zm[r].x;
            const double vzp = c ? zp[r].y : zp[r].x;
            double rr = cy.x * (vym - tt);
            rr = fma(cy.y, vyp - tt, rr);
            rr = fma(cz.x, vzm - tt, rr);
            rr = fma(cz.y, vzp - tt, rr);
            if (has_src && rb + r < nrows) {
              const int64_t idx = kbase + (int64_t)(j0 + rb + r) * nx + i + c;
              double sv = dense ? dense[idx] : 0.0;
              if (st.n) {
                const uint8_t vv = vol[idx];
#pragma unroll
                for (int s = 0; s < 8; ++s)
                  if (s < st.n && st.idx[s] == vv) sv += st.val[s];
              }
              rr = fma(csrc, sv, rr);
            }
            out[c] = fma(2.0, tt, rr);
          }
          *reinterpret_cast<double2 *>(bcol + (rb + r) * Sr) = make_double2(out[0], out[1]);
        }
      }
    }
    HS2_MARK(0);
    __syncthreads();
    HS2_MARK(1);

    // ------------------------------------------------ phase 2: solve along x
    XTab<M> tb;
    tb.b = xtab + (int64_t)lid * (HS2_T_PLANES * (M / 2)) * P + pc;
    tb.P = P;
    const double *ge = GE + ((int64_t)lid * P + pc) * (2 * P);
    double v[M];
    double yf, last;
    if (full) {
#pragma unroll
      for (int t = 0; t < M; t += 2) {
        const double2 x = *reinterpret_cast<const double2 *>(mine + t);
        v[t] = x.x;
        v[t + 1] = x.y;
      }
      yf = xf_forward_full<M>(v, tb);
      last = v[M - 1];
    } else {
#pragma unroll
      for (int t = 0; t < M; t += 2) {
        double2 x = make_double2(0.0, 0.0);
        if (t < rows) x = *reinterpret_cast<const double2 *>(mine + t);
        v[t] = x.x;
        v[t + 1] = x.y;
      }
      yf = xf_forward_short<M>(v, tb, rows, &last);
    }
    if (p2 < P) {
      Y[(2 * p2) * R + r2] = yf;
      Y[(2 * p2 + 1) * R + r2] = last;
    }
    HS2_MARK(2);
    __syncthreads();
    HS2_MARK(3);
    const double E = chunk_interface(ge, Y, P, R, r2, pc, band);
    if (p2 < P) Es[p2 * R + r2] = E;
    HS2_MARK(4);
    __syncthreads();
    HS2_MARK(5);
    const double alpha = (p2 > 0 && p2 < P) ? Es[(p2 - 1) * R + r2] : 0.0;
    if (full)
      xf_backward_full<M>(v, tb, alpha, E);
    else
      xf_backward_short<M>(v, tb, rows, alpha, E);
    if (p2 < P) {
#pragma unroll
      for (int t = 0; t < M; t += 2)
        if (t < rows) *reinterpret_cast<double2 *>(mine + t) = make_double2(v[t], v[t + 1]);
    }
    HS2_MARK(6);
    __syncthreads();
    HS2_MARK(7);

    // ------------------------------------------------ phase 3: d1 = w - 2T, coalesced store
    double *Wk = Wout + kbase + (int64_t)j0 * nx;
    for (int i = 2 * tid; i < nx; i += 2 * nthreads) {
      const double *bcol = buf + i + PAD * (i / M);
      double2 t0[R], w[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        t0[r] = *reinterpret_cast<const double2 *>(Tk + (ro[r + 1] + i));
#pragma unroll
      for (int r = 0; r < R; ++r) w[r] = *reinterpret_cast<const double2 *>(bcol + r * Sr);
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r < nrows)
          *reinterpret_cast<double2 *>(Wk + (r * nx + i)) =
              make_double2(fma(-2.0, t0[r].x, w[r].x), fma(-2.0, t0[r].y, w[r].y));
    }
    HS2_MARK(8);
    __syncthreads();  // buf, Y, Es are rewritten by the next tile
  }
}

template <int M, int R, typename CID>
int launch_xf(hs2_plan *p, const double *T, double *W, const hs2_source *src, const SrcTab &tabsrc, const double *halo_lo,
              const double *halo_hi, int part, cudaStream_t st) {
  const hs2_plan_desc &d = p->d;
  const hs2_axis_tables &ax = d.axis[0];
  const int P = ax.n_chunks;
  const int threads = R * P;
  const int Sr = xf_row_pitch(P, M, xf_pad(R, M));
  const int coef_in_smem = d.n_classes <= 256 ? 1 : 0;
  const size_t smem = ((size_t)R * Sr + 3 * (size_t)P * R + (coef_in_smem ? (size_t)d.n_classes * HS2_COEF_STRIDE : 0)) *
                      sizeof(double);
  HS2_REQUIRE(smem <= (size_t)p->max_smem_optin, "x sweep: tile needs %zu B of shared memory", smem);
  auto kern = sweep_xf_kernel<M, R, CID>;
  if (smem > 48 * 1024) HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int resident = 65536 / (threads * 128) > 0 ? 65536 / (threads * 128) : 1;
  while (resident > 1 && resident * (smem + 1024) > 227 * 1024) --resident;
  {
    // shared memory for the resident blocks only; the rest of the array stays L1
    // (y neighbours, factor tables and the phase-3 re-read of T)
    static const int carveout_env = getenv("HS2_CARVEOUT_X") ? atoi(getenv("HS2_CARVEOUT_X")) : -1;
    const int carveout =
        carveout_env >= 0 ? carveout_env : (int)((resident * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout > 100 ? 100 : carveout));
  }
  const int tiles_y = (int)((d.ny + R - 1) / R);
  const int64_t n_tiles = d.nz * tiles_y;
  HS2_REQUIRE(n_tiles < ((int64_t)1 << 31), "x sweep: too many tiles");
  const int64_t n_work = part == 0 ? n_tiles : (part == 1 ? n_tiles - 2 * tiles_y : 2 * (int64_t)tiles_y);
  if (n_work <= 0) return HS2_OK;
  int64_t blocks = (int64_t)p->sm_count * resident;
  if (blocks > n_work) blocks = n_work;
  const uint8_t *vol = (src && tabsrc.n) ? src->d_vol_elements : nullptr;
  const double *dense = src ? src->d_dense : nullptr;
  static const int pf = getenv("HS2_X_PREFETCH") ? atoi(getenv("HS2_X_PREFETCH")) : 0;   // measured on B200: the prefetch costs 0.05 ms
  kern<<<(unsigned)blocks, threads, smem, st>>>(T, W, (const CID *)d.d_class_id, d.d_class_coef, d.n_classes, coef_in_smem,
                                                 vol, tabsrc, dense, halo_lo, halo_hi, ax.d_line_id,
                                                 reinterpret_cast<const double2 *>(ax.d_tab_il), ax.d_GE, (int)d.nz,
                                                 (int)d.ny, (int)d.nx, P, ax.band, tiles_y, (int)n_tiles, pf, part);
  HS2_CUDA_CHECK(cudaGetLastError());
  return HS2_OK;
}

template <typename CID>
int dispatch_xf(hs2_plan *p, const double *T, double *W, const hs2_source *src, const SrcTab &tabsrc,
                const double *halo_lo, const double *halo_hi, int part, cudaStream_t st) {
  // 8-line tiles.  4-line tiles (HS2_X_R=4: twice the blocks per SM for lines of
  // more than 16 chunks) measured slower on B200 at nx = 512 and nx = 1024.
  static const int r_env = getenv("HS2_X_R") ? atoi(getenv("HS2_X_R")) : 0;
  const bool small = r_env == 4;
  if (small) {
    switch (p->d.axis[0].chunk) {
      case 8: return launch_xf<8, 4, CID>(p, T, W, src, tabsrc, halo_lo, halo_hi, part, st);
      case 16: return launch_xf<16, 4, CID>(p, T, W, src, tabsrc, halo_lo, halo_hi, part, st);
      case 32: return launch_xf<32, 4, CID>(p, T, W, src, tabsrc, halo_lo, halo_hi, part, st);
    }
  }
  switch (p->d.axis[0].chunk) {
    case 8: return launch_xf<8, 8, CID>(p, T, W, src, tabsrc, halo_lo, halo_hi, part, st);
    case 16: return launch_xf<16, 8, CID>(p, T, W, src, tabsrc, halo_lo, halo_hi, part, st);
    case 32: return launch_xf<32, 8, CID>(p, T, W, src, tabsrc, halo_lo, halo_hi, part, st);
  }
  hs2_set_error("x sweep: unsupported chunk size %d", p->d.axis[0].chunk);
  return HS2_E_INVALID;
}

}  // namespace

bool hs2_tile_xf_supported(const hs2_plan *p) {
  if (!hs2_tile_supported(p, 0)) return false;
  const hs2_plan_desc &d = p->d;
  if (d.nx >= ((int64_t)1 << 30) || d.ny >= ((int64_t)1 << 30) || d.nz >= ((int64_t)1 << 30)) return false;
  if (d.ny * d.nx >= ((int64_t)1 << 31)) return false;   // 32-bit in-plane offsets
  if (d.nx & 1) return false;                            // column pairs need 16-byte aligned rows
  return d.axis[0].n_chunks * 8 <= 256 && d.axis[0].d_tab_il != nullptr;
}

int hs2_tile_sweep_xf(hs2_plan *p, const double *T, double *W, const hs2_source *src, const double *halo_lo,
                      const double *halo_hi, int part, cudaStream_t st) {
  HS2_REQUIRE(part == 0 || (part >= 1 && part <= 2 && p->d.nz >= 2), "x sweep: bad part %d", part);
  SrcTab tabsrc;
  int rc = hs2_make_src_tab(src, &tabsrc);
  if (rc) return rc;
  p->last_kernel[0] = HS2_K_X_FOLD;
  if (p->d.class_id_bytes == 1) return dispatch_xf<uint8_t>(p, T, W, src, tabsrc, halo_lo, halo_hi, part, st);
  return dispatch_xf<uint16_t>(p, T, W, src, tabsrc, halo_lo, halo_hi, part, st);
}

#ifdef HS2_PHASE_TIMING
extern "C" int hs2_debug_phase_xf(unsigned long long *out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_hs2_phase, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_hs2_phase, z, sizeof(z));
  }
  return 0;
}
#endif

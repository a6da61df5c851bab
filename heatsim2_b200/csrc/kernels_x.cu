// kernels_x.cu - stage 0 of the ADI step in one kernel: explicit 7-point right
// hand side + implicit solve along the contiguous axis.
//   d1 = (I - 1/2 M^-1 Lx)^-1  M^-1 [ (Lx + Ly + Lz) T + D s ]
// Replaces B0.dot(T) + Dvec*src + tridiagsolve of stage 0
// (heatsim2/alternatingdirection_c_pyx.pyx:397-412): neither the CSR matrix nor
// the right hand side ever exist in HBM.  HBM traffic per cell: read T once
// (y/z neighbours are re-reads served by L1/L2), write d1 once, + 1-2 B class id.
//
// A block owns R consecutive x-lines (rows j0..j0+R-1 of plane k).
//  phase 1  all threads walk the tile column-wise (coalesced along x), keep the
//           y-window (j-1, j, j+1) in registers, fetch the z neighbours from
//           global memory and write the right hand side into shared memory in
//           a chunk-padded layout;
//  phase 2  thread (r, p) takes chunk p (M consecutive cells) of line r into
//           registers and runs the partitioned tridiagonal solve of
//           chunk_core.cuh (lanes of a warp = different lines, so the per-row
//           factors are warp-uniform loads);
//  phase 3  the solution goes back through shared memory and is stored
//           coalesced.
// Shared-memory layout: cell (r, i) at r*S_r + i + i/M doubles - one pad word
// per chunk and S_r = 2 mod 16 make both the row-major phases and the
// chunk-major phase bank-conflict free for 8-byte accesses.
#include "chunk_core.cuh"
#include "tma_util.cuh"
#include <stdlib.h>

namespace {

constexpr int R = 8;  // lines per block

struct SrcTab {
  int n;
  uint8_t idx[8];
  double val[8];
};

__host__ __device__ inline int row_pitch(int P, int M) {
  int s = P * (M + 1);
  while ((s & 15) != 2) ++s;
  return s;
}

template <int M, typename CID>
__global__ void __launch_bounds__(256, (M >= 32 ? 2 : 4))
sweep_x_kernel(const double *__restrict__ T, double *__restrict__ Wout, const CID *__restrict__ cid,
               const double *__restrict__ coef_g, int n_classes, int coef_in_smem,
               const uint8_t *__restrict__ vol, SrcTab st, const double *__restrict__ dense,
               const double *__restrict__ halo_lo, const double *__restrict__ halo_hi,
               const uint32_t *__restrict__ line_id, const double *__restrict__ tab, const double *__restrict__ GE,
               int nz, int ny, int nx, int pitch, int P, int band, int tiles_y, int n_tiles, int tab_in_smem, int pf) {
  extern __shared__ double sm[];
  const int Sr = row_pitch(P, M);
  double *buf = sm;                       // [R][Sr]
  double *Y = buf + R * Sr;               // [2P][R]
  double *Es = Y + 2 * P * R;             // [P][R]
  double *cfs = Es + P * R;               // [n_classes][8] when coef_in_smem
  double *s_tab = cfs + (coef_in_smem ? n_classes * HS2_COEF_STRIDE : 0);   // [HS2_T_PLANES][pitch] when tab_in_smem
  double *s_ge = s_tab + HS2_T_PLANES * pitch;                              // [P][2P]
  HS2_MARK_DECL;
  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;
  const int64_t plane = (int64_t)ny * nx;
  // persistent block: class coefficients and the factor tables of the first
  // line's class are copied to shared memory once and serve every tile
  uint32_t lid_c = 0xffffffffu;
  if (coef_in_smem)
    for (int q = tid; q < n_classes * HS2_COEF_STRIDE; q += nthreads) cfs[q] = coef_g[q];
  if (tab_in_smem && (int)blockIdx.x < n_tiles) {
    lid_c = line_id[(int64_t)(blockIdx.x / tiles_y) * ny + (blockIdx.x % tiles_y) * R];
    const double *gt = tab + (int64_t)lid_c * HS2_T_PLANES * pitch;
    for (int e = tid; e < HS2_T_PLANES * pitch; e += nthreads) s_tab[e] = gt[e];
    const double *gg = GE + (int64_t)lid_c * P * 2 * P;
    for (int e = tid; e < 2 * P * P; e += nthreads) s_ge[e] = gg[e];
  }
  __syncthreads();
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  const int k = tile / tiles_y;
  const int j0 = (tile % tiles_y) * R;
  const int64_t kbase = (int64_t)k * plane;
  const double *coef = coef_in_smem ? cfs : coef_g;
  const int nrows = min(R, ny - j0);
  if (pf) {
    // the only rows of a tile nobody has touched yet are its z+ neighbours:
    // ask L2 for those of this block's NEXT tile now (128-byte lines)
    const int tn = tile + gridDim.x;
    if (tn < n_tiles) {
      const int kn = tn / tiles_y, jn = (tn % tiles_y) * R;
      if (kn + 1 < nz) {
        const double *nxt = T + (int64_t)(kn + 1) * plane + (int64_t)jn * nx;
        const int n_el = min(R, ny - jn) * nx;
        for (int e = tid * 16; e < n_el; e += nthreads * 16) prefetch_l2(nxt + e);
      }
    }
  }

  // ------------------------------------------------ phase 1: right hand side
  // z neighbours: the planes below/above, the neighbouring slab's halo plane,
  // or (domain face, conductance 0) the cell itself
  const double *zlo = k > 0 ? T + kbase - plane : (halo_lo ? halo_lo : T + kbase);
  const double *zhi = k < nz - 1 ? T + kbase + plane : (halo_hi ? halo_hi : T + kbase);
  const double *Tk = T + kbase;
  const CID *cidk = cid + kbase;
  const bool has_src = dense != nullptr || st.n > 0;
  constexpr int RH = R / 2;               // rows handled per register batch
  // each thread owns two adjacent columns (i, i+1): 16-byte global accesses,
  // and the x neighbours inside the pair come from registers
  for (int i = 2 * tid; i < nx; i += 2 * nthreads) {
    const int im = i > 0 ? i - 1 : 0;
    const int ip = i + 2 < nx ? i + 2 : nx - 1;
    double *bcol0 = buf + i + i / M;
    double *bcol1 = buf + (i + 1) + (i + 1) / M;
    int last_id = -1;
    double2 cx = make_double2(0, 0), cy = cx, cz = cx;
    double csrc = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int rb = h * RH;
      double2 tcol[RH + 2], zm[RH], zp[RH];
      double xm[RH], xp[RH];
      int id[RH];
      // issue every load of the batch before any arithmetic
#pragma unroll
      for (int r = -1; r <= RH; ++r) {
        int j = j0 + rb + r;
        j = j < 0 ? 0 : (j > ny - 1 ? ny - 1 : j);
        tcol[r + 1] = *reinterpret_cast<const double2 *>(Tk + (int64_t)j * nx + i);
      }
#pragma unroll
      for (int r = 0; r < RH; ++r) {
        const int j = min(j0 + rb + r, ny - 1);
        const int64_t rowo = (int64_t)j * nx;
        xm[r] = Tk[rowo + im];
        xp[r] = Tk[rowo + ip];
        zm[r] = *reinterpret_cast<const double2 *>(zlo + rowo + i);
        zp[r] = *reinterpret_cast<const double2 *>(zhi + rowo + i);
        if (sizeof(CID) == 1)
          id[r] = *reinterpret_cast<const uint16_t *>(cidk + rowo + i);
        else
          id[r] = (int)*reinterpret_cast<const uint32_t *>(cidk + rowo + i);
      }
#pragma unroll
      for (int r = 0; r < RH; ++r) {
        const double2 tc = tcol[r + 1];
        const int id0 = sizeof(CID) == 1 ? (id[r] & 0xff) : (id[r] & 0xffff);
        const int id1 = sizeof(CID) == 1 ? ((id[r] >> 8) & 0xff) : ((id[r] >> 16) & 0xffff);
        double out[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int idc = c ? id1 : id0;
          if (idc != last_id) {             // interior cells share one class: usually not taken
            const double2 *c2 = reinterpret_cast<const double2 *>(coef + idc * HS2_COEF_STRIDE);
            cx = c2[0];
            cy = c2[1];
            cz = c2[2];
            csrc = c2[3].x;
            last_id = idc;
          }
          const double t0 = c ? tc.y : tc.x;
          const double vxm = c ? tc.x : xm[r];
          const double vxp = c ? xp[r] : tc.y;
          const double vym = c ? tcol[r].y : tcol[r].x;
          const double vyp = c ? tcol[r + 2].y : tcol[r + 2].x;
          const double vzm = c ? zm[r].y : zm[r].x;
          const double vzp = c ? zp[r].y : zp[r].x;
          double rr = cx.x * (vxm - t0);
          rr = fma(cx.y, vxp - t0, rr);
          rr = fma(cy.x, vym - t0, rr);
          rr = fma(cy.y, vyp - t0, rr);
          rr = fma(cz.x, vzm - t0, rr);
          rr = fma(cz.y, vzp - t0, rr);
          if (has_src && rb + r < nrows) {
            const int64_t idx = kbase + (int64_t)(j0 + rb + r) * nx + i + c;
            double sv = dense ? dense[idx] : 0.0;
            if (st.n) {
              const uint8_t vv = vol[idx];
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (q < st.n && st.idx[q] == vv) sv += st.val[q];
            }
            rr = fma(csrc, sv, rr);
          }
          out[c] = rr;
        }
        bcol0[(rb + r) * Sr] = out[0];
        bcol1[(rb + r) * Sr] = out[1];
      }
    }
  }
  HS2_MARK(0);
  __syncthreads();
  HS2_MARK(1);

  // ------------------------------------------------ phase 2: solve along x
  const int r = tid % R;
  const int p = tid / R;
  const bool live = r < nrows && p < P;
  const int pc = p < P ? p : P - 1;
  const int c0 = pc * M;
  const int rows = min(M, nx - c0);
  const bool full = rows == M;
  const int64_t line = (int64_t)k * ny + j0 + (r < nrows ? r : 0);
  const uint32_t lid = line_id[line];
  const double *tb = lid == lid_c ? s_tab + c0 : tab + ((int64_t)lid * HS2_T_PLANES) * pitch + c0;
  const double *ge = lid == lid_c ? s_ge + pc * (2 * P) : GE + ((int64_t)lid * P + pc) * (2 * P);
  double *mine = buf + r * Sr + pc * (M + 1);
  double v[M];
  double yf, last;
  if (full) {
#pragma unroll
    for (int t = 0; t < M; ++t) v[t] = mine[t];
    yf = chunk_forward_full<M>(v, tb, pitch);
    last = v[M - 1];
  } else {
#pragma unroll
    for (int t = 0; t < M; ++t) v[t] = t < rows ? mine[t] : 0.0;
    yf = chunk_forward_short<M>(v, tb, pitch, rows, &last);
  }
  if (p < P) {
    Y[(2 * p) * R + r] = yf;
    Y[(2 * p + 1) * R + r] = last;
  }
  HS2_MARK(2);
  __syncthreads();
  HS2_MARK(3);
  const double E = chunk_interface(ge, Y, P, R, r, pc, band);
  if (p < P) Es[p * R + r] = E;
  HS2_MARK(4);
  __syncthreads();
  HS2_MARK(5);
  const double alpha = (p > 0 && p < P) ? Es[(p - 1) * R + r] : 0.0;
  if (full)
    chunk_backward_full<M>(v, tb, pitch, alpha, E);
  else
    chunk_backward_short<M>(v, tb, pitch, rows, alpha, E);
  if (live) {
#pragma unroll
    for (int t = 0; t < M; ++t)
      if (t < rows) mine[t] = v[t];
  }
  HS2_MARK(6);
  __syncthreads();
  HS2_MARK(7);

  // ------------------------------------------------ phase 3: coalesced store
  double *Wk = Wout + kbase + (int64_t)j0 * nx;
  for (int i = 2 * tid; i < nx; i += 2 * nthreads) {
    const double *bcol0 = buf + i + i / M;
    const double *bcol1 = buf + (i + 1) + (i + 1) / M;
    double2 o[R];
#pragma unroll
    for (int r = 0; r < R; ++r) o[r] = make_double2(bcol0[r * Sr], bcol1[r * Sr]);
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (r < nrows) *reinterpret_cast<double2 *>(Wk + (int64_t)r * nx + i) = o[r];
  }
  HS2_MARK(8);
  __syncthreads();                        // buf, Y, Es are rewritten by the next tile
  }
}

int make_src_tab(const hs2_source *src, SrcTab *st) {
  st->n = 0;
  if (!src || !src->h_value || !src->d_vol_elements) return HS2_OK;
  for (int v = 0; v < 256; ++v) {
    if (src->h_value[v] != 0.0) {
      HS2_REQUIRE(st->n < 8, "hs2_source: more than 8 active volumetric classes in one step; pass a dense source array instead");
      st->idx[st->n] = (uint8_t)v;
      st->val[st->n] = src->h_value[v];
      st->n++;
    }
  }
  return HS2_OK;
}

template <int M, typename CID>
int launch_x(hs2_plan *p, const double *T, double *W, const hs2_source *src, const SrcTab &tabsrc, const double *halo_lo,
             const double *halo_hi, cudaStream_t st) {
  const hs2_plan_desc &d = p->d;
  const hs2_axis_tables &ax = d.axis[0];
  const int P = ax.n_chunks;
  const int threads = R * P;
  const int Sr = row_pitch(P, M);
  const int coef_in_smem = d.n_classes <= 256 ? 1 : 0;
  const size_t base_smem = ((size_t)R * Sr + 3 * (size_t)P * R + (coef_in_smem ? (size_t)d.n_classes * HS2_COEF_STRIDE : 0)) *
                           sizeof(double);
  const size_t tab_smem = ((size_t)HS2_T_PLANES * ax.pitch + 2 * (size_t)P * P) * sizeof(double);
  static const int tabs_env = getenv("HS2_X_TABS_SMEM") ? atoi(getenv("HS2_X_TABS_SMEM")) : 0;   // measured: 4 resident blocks beat 3 with tables in smem
  const int tab_in_smem = (tabs_env && base_smem + tab_smem <= 72 * 1024) ? 1 : 0;
  const size_t smem = base_smem + (tab_in_smem ? tab_smem : 0);
  HS2_REQUIRE(smem <= (size_t)p->max_smem_optin, "x sweep: tile needs %zu B of shared memory", smem);
  auto kern = sweep_x_kernel<M, CID>;
  if (smem > 48 * 1024) HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int resident = 65536 / (threads * 128) > 0 ? 65536 / (threads * 128) : 1;
  while (resident > 1 && resident * (smem + 1024) > 227 * 1024) --resident;
  {
    // shared memory for the resident blocks only; the rest of the array stays L1 (y/x neighbour re-reads)
    static const int carveout_env = getenv("HS2_CARVEOUT_X") ? atoi(getenv("HS2_CARVEOUT_X")) : -1;
    const int carveout = carveout_env >= 0 ? carveout_env : (int)((resident * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout > 100 ? 100 : carveout));
  }
  const int tiles_y = (int)((d.ny + R - 1) / R);
  const int64_t n_tiles = d.nz * tiles_y;
  HS2_REQUIRE(n_tiles < ((int64_t)1 << 31), "x sweep: too many tiles");
  int64_t blocks = (int64_t)p->sm_count * resident;
  if (blocks > n_tiles) blocks = n_tiles;
  const uint8_t *vol = (src && tabsrc.n) ? src->d_vol_elements : nullptr;
  const double *dense = src ? src->d_dense : nullptr;
  static const int pf = getenv("HS2_X_PREFETCH") ? atoi(getenv("HS2_X_PREFETCH")) : 0;   // measured on B200: the prefetch costs 0.05 ms
  kern<<<(unsigned)blocks, threads, smem, st>>>(T, W, (const CID *)d.d_class_id, d.d_class_coef, d.n_classes, coef_in_smem,
                                                 vol, tabsrc, dense, halo_lo, halo_hi, ax.d_line_id, ax.d_tab, ax.d_GE,
                                                 (int)d.nz, (int)d.ny, (int)d.nx, ax.pitch, P, ax.band, tiles_y, (int)n_tiles,
                                                 tab_in_smem, pf);
  HS2_CUDA_CHECK(cudaGetLastError());
  return HS2_OK;
}

template <typename CID>
int dispatch_x(hs2_plan *p, const double *T, double *W, const hs2_source *src, const SrcTab &tabsrc,
               const double *halo_lo, const double *halo_hi, cudaStream_t st) {
  switch (p->d.axis[0].chunk) {
    case 8: return launch_x<8, CID>(p, T, W, src, tabsrc, halo_lo, halo_hi, st);
    case 16: return launch_x<16, CID>(p, T, W, src, tabsrc, halo_lo, halo_hi, st);
    case 32: return launch_x<32, CID>(p, T, W, src, tabsrc, halo_lo, halo_hi, st);
  }
  hs2_set_error("x sweep: unsupported chunk size %d", p->d.axis[0].chunk);
  return HS2_E_INVALID;
}

}  // namespace

bool hs2_tile_x_supported(const hs2_plan *p) {
  if (!hs2_tile_supported(p, 0)) return false;
  const hs2_plan_desc &d = p->d;
  if (d.nx >= ((int64_t)1 << 30) || d.ny >= ((int64_t)1 << 30) || d.nz >= ((int64_t)1 << 30)) return false;
  if (d.nx & 1) return false;             // column pairs need 16-byte aligned rows
  return d.axis[0].n_chunks * R <= 256;
}

int hs2_tile_sweep_x(hs2_plan *p, const double *T, double *W, const hs2_source *src, const double *halo_lo,
                     const double *halo_hi, cudaStream_t st) {
  SrcTab tabsrc;
  int rc = make_src_tab(src, &tabsrc);
  if (rc) return rc;
  if (p->d.class_id_bytes == 1) return dispatch_x<uint8_t>(p, T, W, src, tabsrc, halo_lo, halo_hi, st);
  return dispatch_x<uint16_t>(p, T, W, src, tabsrc, halo_lo, halo_hi, st);
}

#ifdef HS2_PHASE_TIMING
extern "C" int hs2_debug_phase(unsigned long long *out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_hs2_phase, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_hs2_phase, z, sizeof(z));
  }
  return 0;
}
#endif

// kernels_xm.cu - x-sweep, z-marching variant (the default on large grids).
//
// Same mathematics and the same phases 2/3 as kernels_x.cu; the difference is
// how the 7-point stencil gets its z neighbours.  kernels_x.cu reads the
// planes k-1 and k+1 from L2 for every tile (3.25 L2 reads per cell, which
// makes the L2->SM path the bottleneck); here a block keeps a strip of
// HS2_XR x-lines and marches along z, holding the planes k-1 and k of its
// strip in shared memory: per step it loads only plane k+1 (as the z+
// neighbour) plus the two y-halo lines of plane k - 1.25 reads per cell.
// Ownership is fixed: a thread owns column pairs (i, i+1) of all HS2_XR lines,
// reads its z- values from the ring slot, and overwrites them with the z+
// values it just loaded; after the step the slots swap roles.
//
//   shared memory: ring[2][R][nxp] | buf[R][Sr] | Y[2P][R] | Es[P][R] | coef
#include "x_common.cuh"
#include <stdlib.h>

namespace {

constexpr int R = HS2_XR;

template <int M, typename CID>
__global__ void __launch_bounds__(256, 2)
sweep_x_march(const double *__restrict__ T, double *__restrict__ Wout, const CID *__restrict__ cid,
              const double *__restrict__ coef_g, int n_classes, const uint8_t *__restrict__ vol, SrcTab st,
              const double *__restrict__ dense, const double *__restrict__ halo_lo, const double *__restrict__ halo_hi,
              const uint32_t *__restrict__ line_id, const double *__restrict__ tab, const double *__restrict__ GE, int nz,
              int ny, int nx, int nxp, int pitch, int P, int band, int strips, int kz, int n_work) {
  extern __shared__ double sm[];
  const int Sr = hs2_row_pitch(P, M);
  double *ring = sm;                        // [2][R][nxp]
  double *buf = ring + 2 * R * nxp;         // [R][Sr]
  double *Y = buf + R * Sr;                 // [2P][R]
  double *Es = Y + 2 * P * R;               // [P][R]
  double *cfs = Es + P * R;                 // [n_classes][8]
  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;
  const int64_t plane = (int64_t)ny * nx;
  const bool has_src = dense != nullptr || st.n > 0;
  for (int q = tid; q < n_classes * HS2_COEF_STRIDE; q += nthreads) cfs[q] = coef_g[q];

  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
    // work item = (z-chunk, strip); strips vary fastest so that neighbouring
    // blocks share their y-halo lines in L2
    const int j0 = (work % strips) * R;
    const int k0 = (work / strips) * kz;
    const int k1 = min(nz, k0 + kz);
    const int nrows = min(R, ny - j0);
    __syncthreads();                        // previous work item is done with the ring
    // ---- prime the ring: slot (k0&1) <- plane k0, the other <- plane k0-1
    for (int i = 2 * tid; i < nx; i += 2 * nthreads) {
      const double *below = k0 > 0 ? T + (int64_t)(k0 - 1) * plane : (halo_lo ? halo_lo : T);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int j = min(j0 + r, ny - 1);
        const int64_t o = (int64_t)j * nx + i;
        *reinterpret_cast<double2 *>(ring + ((k0 & 1) * R + r) * nxp + i) =
            *reinterpret_cast<const double2 *>(T + (int64_t)k0 * plane + o);
        *reinterpret_cast<double2 *>(ring + (((k0 + 1) & 1) * R + r) * nxp + i) =
            *reinterpret_cast<const double2 *>(below + o);
      }
    }
    __syncthreads();

    for (int k = k0; k < k1; ++k) {
      const int64_t kbase = (int64_t)k * plane;
      const double *Tk = T + kbase;
      const double *zhi = k < nz - 1 ? Tk + plane : (halo_hi ? halo_hi : Tk);
      const CID *cidk = cid + kbase;
      double *cen = ring + (k & 1) * R * nxp;          // plane k
      double *old = ring + ((k + 1) & 1) * R * nxp;    // plane k-1, becomes plane k+1
      // ------------------------------------------------ phase 1: right hand side
      for (int i = 2 * tid; i < nx; i += 2 * nthreads) {
        const int im = i > 0 ? i - 1 : 0;
        const int ip = i + 2 < nx ? i + 2 : nx - 1;
        double2 zp[R], ylo, yhi;
        int id[R];
        // global loads first: plane k+1, the two y-halo lines of plane k, class ids
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int j = min(j0 + r, ny - 1);
          const int64_t o = (int64_t)j * nx + i;
          zp[r] = *reinterpret_cast<const double2 *>(zhi + o);
          if (sizeof(CID) == 1)
            id[r] = *reinterpret_cast<const uint16_t *>(cidk + o);
          else
            id[r] = (int)*reinterpret_cast<const uint32_t *>(cidk + o);
        }
        ylo = *reinterpret_cast<const double2 *>(Tk + (int64_t)(j0 > 0 ? j0 - 1 : 0) * nx + i);
        yhi = *reinterpret_cast<const double2 *>(Tk + (int64_t)min(j0 + R, ny - 1) * nx + i);
        double *bcol0 = buf + i + i / M;
        double *bcol1 = buf + (i + 1) + (i + 1) / M;
        int last_id = -1;
        double2 cx = make_double2(0, 0), cy = cx, cz = cx;
        double csrc = 0.0;
        double2 tprev = ylo;
        double2 tc = *reinterpret_cast<const double2 *>(cen + i);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const double2 tnext = r + 1 < R ? *reinterpret_cast<const double2 *>(cen + (r + 1) * nxp + i) : yhi;
          const double2 tup = (r + 1 < nrows) ? tnext : (r + 1 == nrows ? yhi : tc);
          const double xm0 = cen[r * nxp + im];
          const double xp1 = cen[r * nxp + ip];
          const double2 zm = *reinterpret_cast<const double2 *>(old + r * nxp + i);
          const int idp = id[r];
          const int id0 = sizeof(CID) == 1 ? (idp & 0xff) : (idp & 0xffff);
          const int id1 = sizeof(CID) == 1 ? ((idp >> 8) & 0xff) : ((idp >> 16) & 0xffff);
          double out[2];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int idc = c ? id1 : id0;
            if (idc != last_id) {           // interior cells share one class: usually not taken
              const double2 *c2 = reinterpret_cast<const double2 *>(cfs + idc * HS2_COEF_STRIDE);
              cx = c2[0];
              cy = c2[1];
              cz = c2[2];
              csrc = c2[3].x;
              last_id = idc;
            }
            const double t0 = c ? tc.y : tc.x;
            const double vxm = c ? tc.x : xm0;
            const double vxp = c ? xp1 : tc.y;
            const double vym = c ? tprev.y : tprev.x;
            const double vyp = c ? tup.y : tup.x;
            const double vzm = c ? zm.y : zm.x;
            const double vzp = c ? zp[r].y : zp[r].x;
            double rr = cx.x * (vxm - t0);
            rr = fma(cx.y, vxp - t0, rr);
            rr = fma(cy.x, vym - t0, rr);
            rr = fma(cy.y, vyp - t0, rr);
            rr = fma(cz.x, vzm - t0, rr);
            rr = fma(cz.y, vzp - t0, rr);
            if (has_src && r < nrows) {
              const int64_t idx = kbase + (int64_t)(j0 + r) * nx + i + c;
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
          bcol0[r * Sr] = out[0];
          bcol1[r * Sr] = out[1];
          // plane k-1 is no longer needed at this cell: park plane k+1 there
          *reinterpret_cast<double2 *>(old + r * nxp + i) = zp[r];
          tprev = tc;
          tc = tnext;
        }
      }
      __syncthreads();
      // ------------------------------------------------ phases 2 + 3
      const int r = tid % R;
      const uint32_t lid = line_id[(int64_t)k * ny + j0 + (r < nrows ? r : 0)];
      hs2_x_solve_store<M>(buf, Y, Es, Sr, tid, nthreads, nrows, nx, P, band, pitch, lid, 0xffffffffu, nullptr, nullptr,
                           tab, GE, Wout + kbase + (int64_t)j0 * nx);
      __syncthreads();                      // buf/Y/Es and the ring slots are reused by the next plane
    }
  }
}

template <int M, typename CID>
int launch_xm(hs2_plan *p, const double *T, double *W, const hs2_source *src, const SrcTab &tabsrc, const double *halo_lo,
              const double *halo_hi, cudaStream_t st, bool *done) {
  *done = false;
  const hs2_plan_desc &d = p->d;
  const hs2_axis_tables &ax = d.axis[0];
  const int P = ax.n_chunks;
  const int threads = R * P;
  if (threads > 256 || d.n_classes > 256) return HS2_OK;
  const int Sr = hs2_row_pitch(P, M);
  const int nxp = (int)((d.nx + 1) & ~(int64_t)1);
  const size_t smem = ((size_t)2 * R * nxp + (size_t)R * Sr + 3 * (size_t)P * R + (size_t)d.n_classes * HS2_COEF_STRIDE) *
                      sizeof(double);
  if (smem > 110 * 1024) return HS2_OK;
  static const int kz_env = getenv("HS2_X_KZ") ? atoi(getenv("HS2_X_KZ")) : 32;
  const int kz = kz_env > 0 ? kz_env : 32;
  if (d.nz < 4) return HS2_OK;              // nothing to re-use
  auto kern = sweep_x_march<M, CID>;
  HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int strips = (int)((d.ny + R - 1) / R);
  const int zchunks = (int)((d.nz + kz - 1) / kz);
  const int64_t n_work = (int64_t)strips * zchunks;
  if (n_work >= ((int64_t)1 << 31)) return HS2_OK;
  int64_t blocks = (int64_t)p->sm_count * 2;
  if (blocks > n_work) blocks = n_work;
  const uint8_t *vol = (src && tabsrc.n) ? src->d_vol_elements : nullptr;
  const double *dense = src ? src->d_dense : nullptr;
  kern<<<(unsigned)blocks, threads, smem, st>>>(T, W, (const CID *)d.d_class_id, d.d_class_coef, d.n_classes, vol, tabsrc,
                                                 dense, halo_lo, halo_hi, ax.d_line_id, ax.d_tab, ax.d_GE, (int)d.nz,
                                                 (int)d.ny, (int)d.nx, nxp, ax.pitch, P, ax.band, strips, kz, (int)n_work);
  HS2_CUDA_CHECK(cudaGetLastError());
  *done = true;
  return HS2_OK;
}

}  // namespace

// returns HS2_OK with *done == false when this variant does not apply
int hs2_tile_sweep_x_march(hs2_plan *p, const double *T, double *W, const hs2_source *src, const double *halo_lo,
                           const double *halo_hi, cudaStream_t st, bool *done) {
  *done = false;
  // measured on B200 (profiles/NOTES_r01.md): halves the L2 traffic but runs at 2 blocks/SM and is slower than
  // the tile-per-step kernel, so it is opt-in (HS2_X_MARCH=1) until its issue efficiency is fixed
  static const bool off = !(getenv("HS2_X_MARCH") != nullptr && getenv("HS2_X_MARCH")[0] == '1');
  if (off) return HS2_OK;
  SrcTab tabsrc;
  int rc = hs2_make_src_tab(src, &tabsrc);
  if (rc) return rc;
  const int M = p->d.axis[0].chunk;
  if (p->d.class_id_bytes != 1) return HS2_OK;
  switch (M) {
    case 8: return launch_xm<8, uint8_t>(p, T, W, src, tabsrc, halo_lo, halo_hi, st, done);
    case 16: return launch_xm<16, uint8_t>(p, T, W, src, tabsrc, halo_lo, halo_hi, st, done);
    case 32: return launch_xm<32, uint8_t>(p, T, W, src, tabsrc, halo_lo, halo_hi, st, done);
  }
  return HS2_OK;
}

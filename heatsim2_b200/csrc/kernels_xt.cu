// kernels_xt.cu - stage 0 of the ADI step (x sweep) fed by the tensor copy engine.
//
//   A d1 = M^-1 [ (Lx + Ly + Lz) T + D s ] ,   A = I - 1/2 M^-1 Lx        (delta form)
//
// replaces B0.dot(T) + Dvec*src + tridiagsolve of the reference's stage 0
// (heatsim2/alternatingdirection_c_pyx.pyx:397-412, heatsim2/tridiag.pyx:46-69).
//
// Why a second x kernel (profiles/NOTES_r02.md): ncu of the folded kernel
// (kernels_xf.cu) showed the L1 data pipe at 67 % of peak with 16 warps per SM
// waiting on global loads - the 7 stencil streams go global -> registers -> shared
// (transposition) -> registers, and the T re-read and the coalesced store go
// through the LSU again.  Here the field reaches the SM only through TMA:
//
//  * a block owns a PATCH of 8 x-lines = 2 planes x 4 rows.  Everything the patch
//    needs - 2 x 6 rows of the two planes (y halo included) and 4 rows each of the
//    planes below and above - arrives as three bulk tensor copies (UTMALDG, rank-4
//    map, 128-byte swizzle) = 20 rows of smem per 8 lines;
//  * the tensor map views a row of nx doubles as nx/16 segments of 16:
//    dims (16, ny, nz, nx/16), strides (nx*8, ny*nx*8, 128) bytes.  A box
//    (16, rows, planes, nx/16) lands as [segment][plane][row][16 doubles]; 16-byte
//    unit u of 128-byte line L sits at L*128 + ((u ^ (L & 7)) << 4) (128-byte
//    swizzle), so a quarter warp (4 rows x 2 adjacent segments) reads eight
//    different 16-byte bank groups -> the CHUNK-layout reads (thread = 16
//    consecutive cells of one line = one segment) are conflict free and no
//    transposition pass exists (probe: profiles/micro/tma4d_probe.cu);
//  * thread (line, chunk) builds the right hand side of its 16 cells straight from
//    the five streams (x neighbours are its own registers), runs the partitioned
//    solve of chunk_core.cuh in registers, writes d1 back into the (dead) buffers of
//    the lower/upper plane in the same swizzled layout, and one thread issues the
//    bulk tensor store (UTMASTG).  The next patch's loads are issued as soon as the
//    stage has been read, so they travel during the solve.
//
// Measured ceiling of this data movement alone (tma4d_probe mode 2, 512^3, 2 blocks
// per SM): 0.43 ms against 0.83 ms for the whole folded kernel.
//
// Applicability: nx a multiple of 16, nx <= 512, x-axis tables built for chunk 16.
// Everything else (odd sizes, longer lines) stays on kernels_xf.cu / kernels_v1.cu.
#include "x_common.cuh"
#include "tma_util.cuh"
#include <stdlib.h>

namespace {

constexpr int XT_M = 16;      // cells per chunk = one 128-byte segment
constexpr int XT_LINES = 8;   // lines per patch: 2 planes x 4 rows

struct XtRanges {             // plane ranges [k0, k1) this launch covers (hs2_sweep_x_part)
  int n;
  int k0[2], k1[2];
};

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

template <typename CID>
__device__ __forceinline__ int xt_cell_class(const uint32_t *idw, int e) {
  return sizeof(CID) == 1 ? (int)((idw[e >> 2] >> (8 * (e & 3))) & 0xff) : (int)((idw[e >> 1] >> (16 * (e & 1))) & 0xffff);
}

// all 16 cells of the chunk in one equation class?
template <typename CID>
__device__ __forceinline__ bool xt_one_class(const uint32_t *idw) {
  if (sizeof(CID) == 1) {
    const uint32_t pat = (idw[0] & 0xffu) * 0x01010101u;
    return idw[0] == pat && idw[1] == pat && idw[2] == pat && idw[3] == pat;
  }
  const uint32_t pat = (idw[0] & 0xffffu) * 0x00010001u;
  bool ok = true;
#pragma unroll
  for (int q = 0; q < 8; ++q) ok = ok && idw[q] == pat;
  return ok;
}

struct XtMaps {
  CUtensorMap C;     // T: box (16, 6, 2, S)   two planes with y halo
  CUtensorMap Z;     // T: box (16, 4, 1, S)   one plane, the patch rows
  CUtensorMap Hlo;   // halo_lo plane (nz = 1): box (16, 4, 1, S); unused when no halo
  CUtensorMap Hhi;   // halo_hi plane
  CUtensorMap O;     // work: box (16, 4, 1, S)
};

// patch index -> (range, first plane, first row)
__device__ __forceinline__ void xt_decode(int t, const XtRanges &rg, int tiles_y, int *range, int *k0, int *j0) {
  int r = 0;
  int base = 0;
  const int n0 = ((rg.k1[0] - rg.k0[0] + 1) >> 1) * tiles_y;
  if (rg.n > 1 && t >= n0) {
    r = 1;
    base = n0;
  }
  const int q = t - base;
  *range = r;
  *k0 = rg.k0[r] + 2 * (q / tiles_y);
  *j0 = 4 * (q % tiles_y);
}

template <typename CID, int TABSRC>
__global__ void __launch_bounds__(256, 2)
sweep_xt_kernel(const __grid_constant__ XtMaps tm, const __grid_constant__ UTab ut, const uint8_t *__restrict__ ucode,
                double *__restrict__ Wout, const CID *__restrict__ cid, const double *__restrict__ coef_g, int n_classes, int coef_in_smem,
                const uint8_t *__restrict__ vol, SrcTab st, const double *__restrict__ dense, int has_halo_lo,
                int has_halo_hi, const uint32_t *__restrict__ line_id, const double *__restrict__ tab, int pitch,
                const double *__restrict__ GE, int nz, int ny, int nx, int P, int band, int tiles_y, int n_tiles, XtRanges rg,
                uint32_t szC, uint32_t szZ, int ge_w) {
  extern __shared__ __align__(1024) unsigned char xsm[];
  unsigned char *sC = xsm;
  unsigned char *sZL = sC + szC;
  unsigned char *sZH = sZL + szZ;
  double *Y = reinterpret_cast<double *>(sZH + szZ);   // [2P][8]
  double *Es = Y + 2 * P * XT_LINES;                    // [P][8]
  uint64_t *bar = reinterpret_cast<uint64_t *>(Es + P * XT_LINES);   // [0] C box, [1] ZL + ZH boxes
  double2 *s_ge = reinterpret_cast<double2 *>(bar + 2);               // [P][ge_w] compact interface rows of one line
  double *s_edge = reinterpret_cast<double *>(s_ge + P * ge_w);        // [2][HS2_T_PLANES][16] first / last chunk tables
  double *cfs = s_edge + 3 * HS2_T_PLANES * XT_M;                       // [n_classes][8] when coef_in_smem
  const int tid = threadIdx.x;
  const int lane = tid & 31, wrp = tid >> 5;
  const int rw = lane & 3, pc = (lane >> 2) & 1, pz = (lane >> 3) & 1, pp = lane >> 4;
  const int p = 4 * wrp + 2 * pp + pc;     // chunk = segment of this thread
  const int ln = pz * 4 + rw;              // line within the patch
  const bool active = p < P;
  const int64_t plane = (int64_t)ny * nx;
  // bytes the copy engine delivers per box (the buffers szC / szZ are rounded up to the 1024-byte swizzle atom)
  const uint32_t bytesC = (uint32_t)P * 12u * 128u, bytesZ = (uint32_t)P * 4u * 128u;

  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    fence_mbar_init();
  }
  if (coef_in_smem)
    for (int q = tid; q < n_classes * HS2_COEF_STRIDE; q += blockDim.x) cfs[q] = coef_g[q];
  __syncthreads();
  const double *coef = coef_in_smem ? cfs : coef_g;
  const bool has_src = dense != nullptr || st.n > 0;

  // loads of one patch.  ZL = plane k0-1 (below the grid: the lower slab's halo plane, or zeros at a
  // domain face where the conductance is 0); ZH = plane min(k0+2, nz) (nz = the upper halo / zeros)
  auto issue_C = [&](int k0, int j0) {
    mbar_expect_tx(bar, bytesC);
    tma_load_4d(sC, &tm.C, bar, 0, j0 - 1, k0, 0);
  };
  auto issue_Z = [&](int k0, int j0) {
    mbar_expect_tx(bar + 1, 2 * bytesZ);
    if (k0 - 1 < 0 && has_halo_lo)
      tma_load_4d(sZL, &tm.Hlo, bar + 1, 0, j0, 0, 0);
    else
      tma_load_4d(sZL, &tm.Z, bar + 1, 0, j0, k0 - 1, 0);
    const int kh = k0 + 2 < nz ? k0 + 2 : nz;
    if (kh >= nz && has_halo_hi)
      tma_load_4d(sZH, &tm.Hhi, bar + 1, 0, j0, 0, 0);
    else
      tma_load_4d(sZH, &tm.Z, bar + 1, 0, j0, kh, 0);
  };

  int t = blockIdx.x;
  uint32_t lid_c = 0xffffffffu;      // unique line whose interface rows / end-chunk tables sit in shared memory
  if (t < n_tiles) {
    int r, k0, j0;
    xt_decode(t, rg, tiles_y, &r, &k0, &j0);
    if (tid == 0) {
      issue_C(k0, j0);
      issue_Z(k0, j0);
    }
    lid_c = __ldg(line_id + (int64_t)k0 * ny + j0);
    // rows p-band .. p+band of the inverse interface operator, compact: s_ge[p][q - (p - band)]
    const double2 *gg = reinterpret_cast<const double2 *>(GE + (int64_t)lid_c * P * 2 * P);
    for (int e = tid; e < P * ge_w; e += blockDim.x) {
      const int pr = e / ge_w, q = pr - band + e % ge_w;
      s_ge[e] = (q >= 0 && q < P) ? gg[pr * P + q] : make_double2(0.0, 0.0);
    }
    // factor tables of the first and the last chunk (the only ones that differ from the common table on a
    // line of constant coefficients) and, third, the common table itself: [type][plane][16]
    const double *gt = tab + (int64_t)lid_c * HS2_T_PLANES * pitch;
    for (int e = tid; e < 3 * HS2_T_PLANES * XT_M; e += blockDim.x) {
      const int ty = e / (HS2_T_PLANES * XT_M), pl = (e / XT_M) % HS2_T_PLANES, tt = e % XT_M;
      s_edge[e] = ty == 2 ? ut.v[pl][tt] : gt[pl * pitch + (ty ? (P - 1) * XT_M : 0) + tt];
    }
  }
  __syncthreads();
  uint32_t parity = 0;
  // swizzled line numbers of this thread's streams (fixed for the whole launch)
  const int lc = p * 12 + pz * 6 + rw + 1;       // centre row in the C box
  const int lz = p * 4 + rw;                     // row in a Z box
  const unsigned char *rowC = sC + lc * 128, *rowYm = rowC - 128, *rowYp = rowC + 128;
  const int kC = lc & 7, kYm = (lc - 1) & 7, kYp = (lc + 1) & 7;
  // z neighbours: the other plane of the patch lives in the C box ("early" stream A), the plane outside the
  // patch in a Z box ("late" stream B, read after everything else so that its copy may still be travelling)
  const unsigned char *rowA = pz == 0 ? rowC + 6 * 128 : rowC - 6 * 128;
  const int kA = (pz == 0 ? lc + 6 : lc - 6) & 7;
  const unsigned char *rowB = (pz == 0 ? sZL : sZH) + lz * 128;
  const int kB = lz & 7;

  // class ids of the thread's 16 cells and the unique-line id of its line are fetched ONE PATCH AHEAD (plain
  // global loads that stay in flight during a whole patch), so that nothing waits for them
  constexpr int NIDW = sizeof(CID) == 1 ? 4 : 8;
  auto fetch_ids = [&](int rr, int kk0, int jj0, uint32_t (&ids)[NIDW], uint32_t &lid_out) {
    const int kq = kk0 + pz, jq = jj0 + rw;
    lid_out = 0;
    if (active && jq < ny && kq < rg.k1[rr]) {
      const uint4 *q4 = reinterpret_cast<const uint4 *>(cid + ((int64_t)kq * plane + (int64_t)jq * nx + p * XT_M));
      const uint4 a = __ldg(q4);
      ids[0] = a.x, ids[1] = a.y, ids[2] = a.z, ids[3] = a.w;
      if (sizeof(CID) == 2) {
        const uint4 b = __ldg(q4 + 1);
        ids[NIDW - 4] = b.x, ids[NIDW - 3] = b.y, ids[NIDW - 2] = b.z, ids[NIDW - 1] = b.w;
      }
      lid_out = __ldg(line_id + (int64_t)kq * ny + jq);
    } else {
#pragma unroll
      for (int q = 0; q < NIDW; ++q) ids[q] = 0;
    }
  };
  int r = 0, k0 = 0, j0 = 0;
  uint32_t idw[NIDW], idn[NIDW];
  uint32_t lid = 0, lidn = 0;
  if (t < n_tiles) {
    xt_decode(t, rg, tiles_y, &r, &k0, &j0);
    fetch_ids(r, k0, j0, idw, lid);
  }
  bool first = true;
  HS2_MARK_DECL;
  for (; t < n_tiles; t += gridDim.x) {
    const int k = k0 + pz, j = j0 + rw;
    const bool line_ok = active && j < ny && k < rg.k1[r];     // this thread's line exists and belongs to the launch
    const int64_t cell0 = (int64_t)k * plane + (int64_t)j * nx + p * XT_M;
    const bool more = t + (int)gridDim.x < n_tiles;
    int rn = 0, tn_k0 = 0, tn_j0 = 0;
    if (more) {
      xt_decode(t + gridDim.x, rg, tiles_y, &rn, &tn_k0, &tn_j0);
      fetch_ids(rn, tn_k0, tn_j0, idn, lidn);
    }
    // last plane of an odd plane count: the upper plane of the patch does not exist and the lower plane's z+
    // neighbour (the plane above the grid) sits in ZH -> both z streams are "late"
    const bool lone = k0 + 1 >= nz;
    if (tid == 0 && !first) {
      // the previous patch's d1 left through the Z buffers: once the copy engine has read them, fetch this
      // patch's lower / upper plane rows into them (they are consumed last, below)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      issue_Z(k0, j0);
    }
    first = false;
    HS2_MARK(0);
    mbar_wait(bar, parity);
    if (lone) mbar_wait(bar + 1, parity);
    HS2_MARK(1);

    // ------------------------------------------------ right hand side, chunk layout
    double v[XT_M];
    uint8_t uc = 1;
    if (TABSRC >= 1 && ucode != nullptr && active) uc = __ldg(ucode + (int64_t)lid * P + p);   // used after the stencil
    // common case, decided per warp: no source term and every chunk of the warp lies in one equation class ->
    // straight-line code (no per-cell class test), so that the 16 cells' dependency chains interleave
    const bool plain = !has_src && __all_sync(0xffffffffu, xt_one_class<CID>(idw));
    const unsigned char *rA = (lone && pz == 0) ? sZH + lz * 128 : rowA;
    const int keyA = (lone && pz == 0) ? kB : kA;
    if (active) {
      // centre values first: they are also the x neighbours
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double2 c = *reinterpret_cast<const double2 *>(rowC + ((u ^ kC) << 4));
        v[2 * u] = c.x;
        v[2 * u + 1] = c.y;
      }
      // x neighbours across the chunk ends (closed outer faces: conductance 0, any finite value)
      double xl = p > 0 ? *reinterpret_cast<const double *>(rowC - 12 * 128 + (((7 ^ ((lc - 12) & 7)) << 4) + 8)) : v[0];
      const double xr_end = p < P - 1 ? *reinterpret_cast<const double *>(rowC + 12 * 128 + (((lc + 12) & 7) << 4)) : v[XT_M - 1];
      if (plain) {
        const double2 *c2 = reinterpret_cast<const double2 *>(coef + xt_cell_class<CID>(idw, 0) * HS2_COEF_STRIDE);
        const double2 a0 = c2[0], a1 = c2[1], a2 = c2[2];
        const double cxm = a0.x, cxp = a0.y, cym = a1.x, cyp = a1.y;
        const double cA = pz == 0 ? a2.y : a2.x, cB = pz == 0 ? a2.x : a2.y;
        const double csum = -(((cxm + cxp) + (cym + cyp)) + (cA + cB));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double2 ym = *reinterpret_cast<const double2 *>(rowYm + ((u ^ kYm) << 4));
          const double2 yp = *reinterpret_cast<const double2 *>(rowYp + ((u ^ kYp) << 4));
          const double2 za = *reinterpret_cast<const double2 *>(rA + ((u ^ keyA) << 4));
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int e = 2 * u + c;
            const double tt = v[e];
            const double xr = e < XT_M - 1 ? v[e + 1] : xr_end;
            double rr = cxm * (xl - tt);
            rr = fma(cxp, xr - tt, rr);
            rr = fma(cym, (c ? ym.y : ym.x) - tt, rr);
            rr = fma(cyp, (c ? yp.y : yp.x) - tt, rr);
            rr = fma(cA, (c ? za.y : za.x) - tt, rr);
            rr = fma(-cB, tt, rr);             // the late stream adds cB * z below
            xl = tt;
            v[e] = rr;
          }
        }
        (void)csum;
      } else {
        int last_id = -1;
        double cxm = 0, cxp = 0, cym = 0, cyp = 0, cA = 0, cB = 0, csrc = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double2 ym = *reinterpret_cast<const double2 *>(rowYm + ((u ^ kYm) << 4));
          const double2 yp = *reinterpret_cast<const double2 *>(rowYp + ((u ^ kYp) << 4));
          const double2 za = *reinterpret_cast<const double2 *>(rA + ((u ^ keyA) << 4));
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int e = 2 * u + c;
            const int idc = xt_cell_class<CID>(idw, e);
            if (idc != last_id) {
              const double2 *c2 = reinterpret_cast<const double2 *>(coef + idc * HS2_COEF_STRIDE);
              const double2 a0 = c2[0], a1 = c2[1], a2 = c2[2];
              cxm = a0.x, cxp = a0.y, cym = a1.x, cyp = a1.y;
              cA = pz == 0 ? a2.y : a2.x;      // plane 0: z+ is in the patch, plane 1: z-
              cB = pz == 0 ? a2.x : a2.y;
              csrc = c2[3].x;
              last_id = idc;
            }
            const double tt = v[e];
            const double xr = e < XT_M - 1 ? v[e + 1] : xr_end;
            double rr = cxm * (xl - tt);
            rr = fma(cxp, xr - tt, rr);
            rr = fma(cym, (c ? ym.y : ym.x) - tt, rr);
            rr = fma(cyp, (c ? yp.y : yp.x) - tt, rr);
            rr = fma(cA, (c ? za.y : za.x) - tt, rr);
            rr = fma(-cB, tt, rr);
            if (has_src && line_ok) {
              double sv = dense ? dense[cell0 + e] : 0.0;
              if (st.n) {
                const uint8_t vv = vol[cell0 + e];
#pragma unroll
                for (int s = 0; s < 8; ++s)
                  if (s < st.n && st.idx[s] == vv) sv += st.val[s];
              }
              rr = fma(csrc, sv, rr);
            }
            xl = tt;
            v[e] = rr;
          }
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < XT_M; ++e) v[e] = 0.0;
    }
    if (!lone) mbar_wait(bar + 1, parity);
    parity ^= 1;
    if (active) {
      if (plain) {
        const double2 a2 = reinterpret_cast<const double2 *>(coef + xt_cell_class<CID>(idw, 0) * HS2_COEF_STRIDE)[2];
        const double cB = pz == 0 ? a2.x : a2.y;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double2 zb = *reinterpret_cast<const double2 *>(rowB + ((u ^ kB) << 4));
          v[2 * u] = fma(cB, zb.x, v[2 * u]);
          v[2 * u + 1] = fma(cB, zb.y, v[2 * u + 1]);
        }
      } else {
        int last_id = -1;
        double cB = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double2 zb = *reinterpret_cast<const double2 *>(rowB + ((u ^ kB) << 4));
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int e = 2 * u + c;
            const int idc = xt_cell_class<CID>(idw, e);
            if (idc != last_id) {
              const double2 a2 = reinterpret_cast<const double2 *>(coef + idc * HS2_COEF_STRIDE)[2];
              cB = pz == 0 ? a2.x : a2.y;
              last_id = idc;
            }
            v[e] = fma(cB, c ? zb.y : zb.x, v[e]);
          }
        }
      }
    }
    HS2_MARK(2);
    __syncthreads();     // the stage has been read
    HS2_MARK(3);
    if (tid == 0 && more) issue_C(tn_k0, tn_j0);      // travels during the solve

    // ------------------------------------------------ partitioned solve along x (chunk_core.cuh)
    const int pcl = active ? p : 0;
    const bool uni = TABSRC == 1 && ucode != nullptr && __all_sync(0xffffffffu, uc != 0);
    const bool own = lid == lid_c;                                  // tables of this line are (partly) in shared memory
    const bool edge = own && (pcl == 0 || pcl == P - 1);
    const double *tb = edge ? s_edge + (pcl == 0 ? 0 : HS2_T_PLANES * XT_M)
                            : tab + ((int64_t)lid * HS2_T_PLANES) * pitch + pcl * XT_M;
    const int tpitch = edge ? XT_M : pitch;
    // TABSRC 2: every chunk that carries the common table, or is an end chunk of the block's line, reads 8-byte
    // broadcast loads from the three shared-memory tables; a warp takes that path when all its lanes can
    const int ttype = (TABSRC == 2 && ucode != nullptr && uc != 0) ? 2 : (edge ? (pcl == 0 ? 0 : 1) : -1);
    const bool smem_tab = TABSRC == 2 && __all_sync(0xffffffffu, ttype >= 0);
    TabShared ts;
    ts.a = smem_u32(s_edge + (ttype < 0 ? 0 : ttype) * HS2_T_PLANES * XT_M);
    ts.pitch_b = XT_M * 8u;
    double yf, last;
    if (TABSRC == 1 && uni) {
      yf = chunk_forward_const<XT_M>(v, ut);
      last = v[XT_M - 1];
    } else if (smem_tab) {
      yf = chunk_fwd<XT_M, true>(v, ts, XT_M, &last);
    } else {
      yf = chunk_forward_full<XT_M>(v, tb, tpitch);
      last = v[XT_M - 1];
    }
    if (active) {
      Y[(2 * p) * XT_LINES + ln] = yf;
      Y[(2 * p + 1) * XT_LINES + ln] = last;
    }
    HS2_MARK(4);
    __syncthreads();
    HS2_MARK(5);
    double E;
    {
      // E_p = row p of the inverse interface operator times (yf_0, yl_0, yf_1, yl_1, ...), band-limited
      const int q0 = max(0, pcl - band), q1 = min(P - 1, pcl + band);
      const double2 *grow = own ? s_ge + pcl * ge_w - (pcl - band)
                                : reinterpret_cast<const double2 *>(GE + ((int64_t)lid * P + pcl) * (2 * P));
      double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
      int q = q0;
#pragma unroll 2
      for (; q + 1 <= q1; q += 2) {
        const double2 g0 = grow[q], g1 = grow[q + 1];
        e0 = fma(g0.x, Y[(2 * q) * XT_LINES + ln], e0);
        e1 = fma(g0.y, Y[(2 * q + 1) * XT_LINES + ln], e1);
        e2 = fma(g1.x, Y[(2 * q + 2) * XT_LINES + ln], e2);
        e3 = fma(g1.y, Y[(2 * q + 3) * XT_LINES + ln], e3);
      }
      if (q <= q1) {
        const double2 g0 = grow[q];
        e0 = fma(g0.x, Y[(2 * q) * XT_LINES + ln], e0);
        e1 = fma(g0.y, Y[(2 * q + 1) * XT_LINES + ln], e1);
      }
      E = (e0 + e1) + (e2 + e3);
    }
    if (active) Es[p * XT_LINES + ln] = E;
    HS2_MARK(6);
    __syncthreads();
    HS2_MARK(7);
    const double alpha = (active && p > 0) ? Es[(p - 1) * XT_LINES + ln] : 0.0;
    if (TABSRC == 1 && uni)
      chunk_backward_const<XT_M>(v, ut, alpha, E);
    else if (smem_tab)
      chunk_bwd<XT_M, true>(v, ts, XT_M, alpha, E);
    else
      chunk_backward_full<XT_M>(v, tb, tpitch, alpha, E);

    HS2_MARK(8);
    // ------------------------------------------------ d1 -> the dead Z buffers (same swizzle) -> bulk tensor store
    if (active) {
      unsigned char *o = (pz == 0 ? sZL : sZH) + lz * 128;
#pragma unroll
      for (int u = 0; u < 8; ++u) *reinterpret_cast<double2 *>(o + ((u ^ kB) << 4)) = make_double2(v[2 * u], v[2 * u + 1]);
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tma_store_4d(&tm.O, sZL, 0, j0, k0, 0);
      if (k0 + 1 < rg.k1[r]) tma_store_4d(&tm.O, sZH, 0, j0, k0 + 1, 0);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    HS2_MARK(9);
    r = rn, k0 = tn_k0, j0 = tn_j0;
    lid = lidn;
#pragma unroll
    for (int q = 0; q < NIDW; ++q) idw[q] = idn[q];
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

bool encode4(CUtensorMap *m, const void *base, int nx, int ny, int nzz, int rows, int planes) {
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  }
  if (!encode || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
  const int S = nx / 16;
  cuuint64_t dims[4] = {16, (cuuint64_t)ny, (cuuint64_t)nzz, (cuuint64_t)S};
  cuuint64_t strides[3] = {(cuuint64_t)nx * 8, (cuuint64_t)ny * nx * 8, 128};
  cuuint32_t box[4] = {16, (cuuint32_t)rows, (cuuint32_t)planes, (cuuint32_t)S};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void *>(base), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

uint32_t round1k(uint32_t b) { return (b + 1023u) & ~1023u; }

template <typename CID>
int launch_xt(hs2_plan *p, const double *T, double *W, const hs2_source *src, const SrcTab &tabsrc, const double *halo_lo,
              const double *halo_hi, int part, cudaStream_t st, bool *done) {
  *done = false;
  const hs2_plan_desc &d = p->d;
  const hs2_axis_tables &ax = d.axis[0];
  const int nx = (int)d.nx, ny = (int)d.ny, nz = (int)d.nz;
  const int P = ax.n_chunks;
  XtRanges rg;
  if (part == 0) {
    rg.n = 1, rg.k0[0] = 0, rg.k1[0] = nz, rg.k0[1] = rg.k1[1] = 0;
  } else if (part == HS2_X_INTERIOR) {
    rg.n = 1, rg.k0[0] = 1, rg.k1[0] = nz - 1, rg.k0[1] = rg.k1[1] = 0;
  } else {
    rg.n = 2, rg.k0[0] = 0, rg.k1[0] = 1, rg.k0[1] = nz - 1, rg.k1[1] = nz;
  }
  const int tiles_y = (ny + 3) / 4;
  int64_t n_tiles = 0;
  for (int r = 0; r < rg.n; ++r) n_tiles += (int64_t)((rg.k1[r] - rg.k0[r] + 1) / 2) * tiles_y;
  if (n_tiles <= 0) {
    *done = true;
    return HS2_OK;
  }
  if (n_tiles >= ((int64_t)1 << 31)) return HS2_OK;
  XtMaps tm;
  memset(&tm, 0, sizeof(tm));
  if (!encode4(&tm.C, T, nx, ny, nz, 6, 2) || !encode4(&tm.Z, T, nx, ny, nz, 4, 1) || !encode4(&tm.O, W, nx, ny, nz, 4, 1))
    return HS2_OK;
  if (halo_lo && !encode4(&tm.Hlo, halo_lo, nx, ny, 1, 4, 1)) return HS2_OK;
  if (halo_hi && !encode4(&tm.Hhi, halo_hi, nx, ny, 1, 4, 1)) return HS2_OK;
  const uint32_t szC = round1k((uint32_t)P * 12 * 128), szZ = round1k((uint32_t)P * 4 * 128);
  const int coef_in_smem = d.n_classes <= 64 ? 1 : 0;
  const int ge_w = 2 * ax.band + 1;
  const size_t smem = (size_t)szC + 2 * (size_t)szZ + 3 * (size_t)P * XT_LINES * sizeof(double) + 16 +
                      (size_t)P * ge_w * 16 + 3 * HS2_T_PLANES * XT_M * sizeof(double) +
                      (coef_in_smem ? (size_t)d.n_classes * HS2_COEF_STRIDE * sizeof(double) : 0);
  if (smem + 1024 > (size_t)p->max_smem_optin) return HS2_OK;
  const int threads = ((P + 3) / 4) * 32;
  // factor tables of the solve: 0 = global loads, 1 = constant bank for warps on the common table,
  // 2 = three shared-memory tables (common / first chunk / last chunk) read with 8-byte broadcast loads
  const int tabsrc_env = getenv("HS2_XT_TABSRC") ? atoi(getenv("HS2_XT_TABSRC")) : 2;
  const uint8_t *ucode = (p->has_utab[0] && !(d.flags & HS2_FLAG_NO_UTAB)) ? ax.d_ucode : nullptr;
  const int tsrc = ucode ? tabsrc_env : 0;
  auto kern = tsrc == 1 ? sweep_xt_kernel<CID, 1> : (tsrc == 2 ? sweep_xt_kernel<CID, 2> : sweep_xt_kernel<CID, 0>);
  HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int resident = (int)((size_t)227 * 1024 / (smem + 1024));
  const int by_threads = 2048 / threads, by_regs = 65536 / (threads * 128);
  if (resident > by_threads) resident = by_threads;
  if (resident > by_regs) resident = by_regs;
  if (resident < 1) resident = 1;
  static const int res_env = getenv("HS2_XT_RESIDENT") ? atoi(getenv("HS2_XT_RESIDENT")) : 0;
  if (res_env > 0) resident = res_env;
  const int carve = (int)((resident * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
  HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve > 100 ? 100 : carve));
  int64_t blocks = (int64_t)p->sm_count * resident;
  if (blocks > n_tiles) blocks = n_tiles;
  const uint8_t *vol = (src && tabsrc.n) ? src->d_vol_elements : nullptr;
  const double *dense = src ? src->d_dense : nullptr;
  kern<<<(unsigned)blocks, threads, smem, st>>>(tm, p->utab[0], ucode, W, (const CID *)d.d_class_id, d.d_class_coef, d.n_classes,
                                                 coef_in_smem, vol, tabsrc, dense, halo_lo ? 1 : 0, halo_hi ? 1 : 0,
                                                 ax.d_line_id, ax.d_tab, ax.pitch, ax.d_GE, nz, ny, nx, P, ax.band, tiles_y,
                                                 (int)n_tiles, rg, szC, szZ, ge_w);
  HS2_CUDA_CHECK(cudaGetLastError());
  p->last_kernel[0] = HS2_K_X_TMA;
  *done = true;
  return HS2_OK;
}

}  // namespace

bool hs2_tile_xt_supported(const hs2_plan *p) {
  const hs2_plan_desc &d = p->d;
  const hs2_axis_tables &ax = d.axis[0];
  if (d.flags & (HS2_FLAG_FORCE_FALLBACK | HS2_FLAG_X_FOLD)) return false;
  if (ax.chunk != XT_M || !ax.d_tab || !ax.d_GE || ax.pitch <= 0) return false;
  if (d.nx % 16 != 0 || d.nx < 16 || d.nx > 512 || ax.n_chunks != d.nx / 16) return false;
  if (d.ny >= ((int64_t)1 << 30) || d.nz >= ((int64_t)1 << 30)) return false;
  if ((reinterpret_cast<uintptr_t>(d.d_class_id) & 15)) return false;
  return true;
}

int hs2_tile_sweep_xt(hs2_plan *p, const double *T, double *W, const hs2_source *src, const double *halo_lo,
                      const double *halo_hi, int part, cudaStream_t st, bool *done) {
  *done = false;
  SrcTab tabsrc;
  int rc = hs2_make_src_tab(src, &tabsrc);
  if (rc) return rc;
  if (p->d.class_id_bytes == 1) return launch_xt<uint8_t>(p, T, W, src, tabsrc, halo_lo, halo_hi, part, st, done);
  return launch_xt<uint16_t>(p, T, W, src, tabsrc, halo_lo, halo_hi, part, st, done);
}

#ifdef HS2_PHASE_TIMING
extern "C" int hs2_debug_phase_xt(unsigned long long *out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_hs2_phase, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_hs2_phase, z, sizeof(z));
  }
  return 0;
}
#endif

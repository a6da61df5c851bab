// kernels_xw.cu - stage 0 of the ADI step (x sweep), ONE WARP PER LINE, lane = chunk.
//
//   A d1 = M^-1 [ (Lx + Ly + Lz) T + D s ] ,   A = I - 1/2 M^-1 Lx        (delta form)
//
// replaces B0.dot(T) + Dvec*src + tridiagsolve of the reference's stage 0
// (heatsim2/alternatingdirection_c_pyx.pyx:397-412, heatsim2/tridiag.pyx:46-69).
//
// What the profile of the patch kernel (kernels_xt.cu, thread = (line, chunk), 4 chunks x 8 lines per
// warp) showed (profiles/NOTES_r02.md): 1035 instructions per thread per patch of which 302 are FP64,
// 4 block-wide barriers per patch (17 % of the stall samples at the first one alone), the interface
// sums through shared memory (0.4 of its 0.9 L1 wavefronts per cell).  Here a warp owns a whole line:
//
//  * lane p holds chunk p (16 cells = one 128-byte segment) of the line in registers; lines of
//    nx = 512 cells exactly fill a warp.  The chunk interfaces are exchanged with warp shuffles, the
//    x neighbours across chunk ends too: there is NO block-wide barrier in the main loop;
//  * a GROUP of 10 warps owns a patch of 2 planes x 5 rows.  The field arrives through the tensor copy
//    engine: per plane one box of 7 rows (y halo included, rank-4 map, 128-byte swizzle) shared by the
//    group, and per warp ONE private row of the plane below / above the patch.  Box rows per segment
//    are odd (7 and 1), so that the swizzle makes the chunk-layout reads (lane = segment, stride
//    7 x 128 B or 128 B) conflict free.  24 rows of shared memory per 10 lines; two groups per block,
//    one block per SM -> 20 warps per SM;
//  * the last warp of a group to finish reading the shared boxes issues the loads of the group's next
//    patch (shared-memory counter, nobody waits); a warp writes d1 over its private row and stores it
//    with its own bulk tensor store; the only waits are for data (mbarriers);
//  * lines with constant coefficients and closed ends ("ghost-uniform", all lines of the BASELINE
//    grids) use ONE chunk table for all 32 chunks, read with 8-byte broadcast loads from shared memory:
//    the first and the last chunk are the common chunk with a mirrored ghost neighbour (alpha = x_first,
//    beta = x_last), which only changes the interface operator (plan.py ghost_uniform_tables) and adds
//    one correction pass for lane 0.  Other lines read per-lane tables from global memory;
//  * equation classes: cells 1..14 of a chunk take the class of cell 8, the end cells their own
//    (domain faces); chunks with a class change inside take a per-cell path.
//
// Applicability: nx = 512 (32 chunks of 16), no volumetric source in this step (those steps run
// kernels_xt.cu), chunk tables built for chunk 16.
#include "x_common.cuh"
#include "tma_util.cuh"
#include <stdlib.h>

namespace {

constexpr int XW_M = 16;       // cells per chunk
// template parameters of the kernel: XW_R rows per plane per patch (odd: see above), XW_G groups per block;
// a group is 2 * XW_R warps and 2 * xw_box_rows(XW_R) + 2 * XW_R rows of shared memory
constexpr int XW_HDR = 128;    // doubles per unique line in front of its interface rows

struct XwRanges {              // plane ranges [k0, k1) this launch covers (hs2_sweep_x_part)
  int n;
  int k0[2], k1[2];
};

struct XwMaps {
  CUtensorMap C;      // T: box (16, 7, 1, S)
  CUtensorMap Z;      // T: box (16, 1, 1, S)
  CUtensorMap HloZ;   // halo_lo plane (nz = 1): box (16, 1, 1, S)
  CUtensorMap HhiC;   // halo_hi plane: box (16, 7, 1, S)
  CUtensorMap HhiZ;   // halo_hi plane: box (16, 1, 1, S)
  CUtensorMap O;      // work: box (16, 1, 1, S)
};

// (all shared-memory operands are 32-bit shared-window addresses)
__device__ __forceinline__ void xw_tma_load(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c1, int c2, int c3 = 0) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void xw_tma_store(const CUtensorMap *map, uint32_t src, int c1, int c2, int c3 = 0) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__device__ __forceinline__ double xw_lds1(uint32_t a) {
  double r;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a));
  return r;
}

// barrier of the two warps that share a line (WPL = 2)
__device__ __forceinline__ void xw_pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__device__ __forceinline__ void xw_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void xw_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "XW_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra XW_WAIT_DONE;\n"
      "bra XW_WAIT_LOOP;\n"
      "XW_WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

template <typename CID>
__device__ __forceinline__ int xw_cell_class(const uint32_t *idw, int e) {
  return sizeof(CID) == 1 ? (int)((idw[e >> 2] >> (8 * (e & 3))) & 0xff) : (int)((idw[e >> 1] >> (16 * (e & 1))) & 0xffff);
}

// cells 1..14 of the chunk all in the class of cell 8?
template <typename CID>
__device__ __forceinline__ bool xw_interior_one_class(const uint32_t *idw) {
  if (sizeof(CID) == 1) {
    const uint32_t pat = ((idw[2]) & 0xffu) * 0x01010101u;
    return idw[1] == pat && idw[2] == pat && ((idw[0] ^ pat) & 0xffffff00u) == 0 && ((idw[3] ^ pat) & 0x00ffffffu) == 0;
  }
  const uint32_t pat = (idw[4] & 0xffffu) * 0x00010001u;
  bool ok = ((idw[0] ^ pat) & 0xffff0000u) == 0 && ((idw[7] ^ pat) & 0x0000ffffu) == 0;
#pragma unroll
  for (int q = 1; q < 7; ++q) ok = ok && idw[q] == pat;
  return ok;
}

// shuffles inside the P lanes that hold one line
template <int P>
__device__ __forceinline__ double xw_shfl_up(double x, int d) { return __shfl_up_sync(0xffffffffu, x, d, P); }
template <int P>
__device__ __forceinline__ double xw_shfl_down(double x, int d) { return __shfl_down_sync(0xffffffffu, x, d, P); }

__device__ __forceinline__ double2 xw_lds(uint32_t a) {
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(a));
  return r;
}

// rows of a plane box: the patch rows and the y halo, plus one unused row where that makes the count odd
// (odd rows per segment: the swizzle then spreads a quarter warp's segments over all banks)
__host__ __device__ constexpr int xw_box_rows(int r) { return (r + 2) | 1; }

// patch index -> (range, first plane, first row)
template <int XW_R>
__device__ __forceinline__ void xw_decode(int t, const XwRanges &rg, int tiles_y, int *range, int *k0, int *j0) {
  int r = 0;
  int base = 0;
  const int n0 = ((rg.k1[0] - rg.k0[0] + 1) >> 1) * tiles_y;
  if (rg.n > 1 && t >= n0) {
    r = 1;
    base = n0;
  }
  const int q = t - base;
  *range = r;
  *k0 = (r ? rg.k0[1] : rg.k0[0]) + 2 * (q / tiles_y);
  *j0 = XW_R * (q % tiles_y);
}

// XW_R rows per plane per patch, XW_G groups per block, LPW lines per warp: 1 (nx = 512, lane = chunk) or
// 2 (nx = 256: the two half warps hold the same row of the patch's two planes), WPL warps per line: 1 or
// 2 (nx = 1024: warp h of a line holds chunks 32 h .. 32 h + 31; what crosses the middle of the line - one x
// neighbour each way, the interface values within the band, E of chunk 31 - goes through shared memory between
// two barriers of the pair)
template <typename CID, int XW_R, int XW_G, int LPW, int WPL>
__global__ void __maxnreg__(XW_G * 2 * XW_R * WPL / LPW <= 16 ? 128 : 96)
sweep_xw_kernel(const __grid_constant__ XwMaps tm, const CID *__restrict__ cid, const double *__restrict__ coef_g, int n_classes,
                int has_halo_lo, int has_halo_hi, const uint32_t *__restrict__ line_id, const double *__restrict__ tab, int pitch,
                const double *__restrict__ GE, int band_g, const double *__restrict__ xw_tab, const uint8_t *__restrict__ xw_code,
                int n_slots, int ge_w, int nz, int ny, int tiles_y, int n_tiles, XwRanges rg) {
  static_assert(LPW == 1 || WPL == 1, "two lines per warp or two warps per line");
  constexpr int PW = 32 / LPW;                                       // chunks of a line in one warp (shuffle width)
  constexpr int P = PW * WPL;                                        // chunks per line
  constexpr int NX = P * XW_M;
  constexpr uint32_t ROW = P * 128;                                  // bytes per row
  constexpr uint32_t PART = ROW / WPL;                               // bytes of a row one warp loads / stores
  constexpr int WPG = 2 * XW_R * WPL / LPW;                          // warps per group
  constexpr int CR = xw_box_rows(XW_R);                              // rows per segment in a plane box (odd)
  constexpr uint32_t XW_CBOX = CR * ROW;                             // one plane's box
  constexpr uint32_t XW_GROUP = 2 * XW_CBOX + 2 * XW_R * ROW;
  extern __shared__ __align__(1024) unsigned char xsm[];
  double *s_hdr = reinterpret_cast<double *>(xsm + XW_G * XW_GROUP);               // [n_slots][XW_HDR]
  double2 *s_ge = reinterpret_cast<double2 *>(s_hdr + n_slots * XW_HDR);            // [n_slots][ge_w][P]
  double *cfs = reinterpret_cast<double *>(s_ge + n_slots * ge_w * P);               // [n_classes][8]
  double2 *s_y = reinterpret_cast<double2 *>(cfs + n_classes * HS2_COEF_STRIDE);    // [XW_G][2 XW_R][P]   (WPL = 2)
  double *s_e = reinterpret_cast<double *>(s_y + (WPL == 2 ? XW_G * 2 * XW_R * P : 0));   // [XW_G][2 XW_R]
  uint64_t *barC = reinterpret_cast<uint64_t *>(s_e + (WPL == 2 ? XW_G * 2 * XW_R : 0));  // [XW_G]
  uint64_t *barZ = barC + XW_G;                                                      // [XW_G * WPG]
  int *cnt = reinterpret_cast<int *>(barZ + XW_G * WPG);                             // [XW_G]
  uint8_t *s_code = reinterpret_cast<uint8_t *>(cnt + XW_G);                         // [n_slots]

  const int tid = threadIdx.x;
  const int lane = tid & 31, wrp = tid >> 5;
  const int g = wrp / WPG, w = wrp % WPG;
  const int lw = WPL == 2 ? w >> 1 : w;                 // LPW = 1: the warp's line within the group
  const int hx = WPL == 2 ? w & 1 : 0;                  // WPL = 2: which half of the line
  const int pz = LPW == 1 ? lw / XW_R : lane >> 4;      // plane of this lane's line within the patch
  const int rw = LPW == 1 ? lw % XW_R : w;              // row of this lane's line within the patch
  const int p = hx * 32 + (lane & (PW - 1));            // chunk
  const int n_groups = gridDim.x * XW_G;

  const uint32_t gb = smem_u32(xsm) + g * XW_GROUP;
  const uint32_t sC0 = gb, sC1 = gb + XW_CBOX;
  const uint32_t sZ = gb + 2 * XW_CBOX + (pz * XW_R + rw) * ROW;    // this line's private row
  const uint32_t bC = smem_u32(barC + g), bZ = smem_u32(barZ + wrp);
  const uint32_t coef_s = smem_u32(cfs);

  if (tid == 0) {
    for (int q = 0; q < XW_G; ++q) {
      mbar_init(barC + q, 1);
      cnt[q] = 0;
    }
    for (int q = 0; q < XW_G * WPG; ++q) mbar_init(barZ + q, 1);
    fence_mbar_init();
  }
  {
    const int64_t stride = XW_HDR + ge_w * P * 2;
    for (int q = tid; q < n_slots * XW_HDR; q += blockDim.x) s_hdr[q] = xw_tab[(int64_t)(q / XW_HDR) * stride + q % XW_HDR];
    for (int q = tid; q < n_slots * ge_w * P; q += blockDim.x) {
      const int s = q / (ge_w * P), e = q % (ge_w * P);
      s_ge[q] = reinterpret_cast<const double2 *>(xw_tab + (int64_t)s * stride + XW_HDR)[e];
    }
    for (int q = tid; q < n_slots; q += blockDim.x) s_code[q] = xw_code[q];
    for (int q = tid; q < n_classes * HS2_COEF_STRIDE; q += blockDim.x) cfs[q] = coef_g[q];
  }
  __syncthreads();

  // loads of the two shared boxes of patch t (planes k0 and k0+1, rows j0-1 .. j0+CR-2); the plane above the
  // grid is the upper slab's halo plane, or zeros at a domain face (conductance 0 there)
  auto issue_C = [&](int t) {
    int r, k0, j0;
    xw_decode<XW_R>(t, rg, tiles_y, &r, &k0, &j0);
    xw_expect_tx(bC, 2 * XW_CBOX);
    xw_tma_load(sC0, &tm.C, bC, j0 - 1, k0);
    if (k0 + 1 >= nz && has_halo_hi)
      xw_tma_load(sC1, &tm.HhiC, bC, j0 - 1, 0);
    else
      xw_tma_load(sC1, &tm.C, bC, j0 - 1, k0 + 1);
  };

  int t = blockIdx.x * XW_G + g;
  if (t < n_tiles && w == 0 && lane == 0) issue_C(t);

  // swizzled addresses of this lane's streams (fixed for the whole launch).  Plane box: [segment p][CR rows][16]:
  // 128-byte line CR * p + r, 16-byte unit u at ((u ^ (line & 7)) << 4); private row: line p
  const int lc = CR * p + rw + 1;
  const uint32_t rowC = (pz == 0 ? sC0 : sC1) + lc * 128;
  const uint32_t rowA = (pz == 0 ? sC1 : sC0) + lc * 128;       // the other plane of the patch (same key)
  const uint32_t kC = (lc & 7) << 4, kYm = ((lc - 1) & 7) << 4, kYp = ((lc + 1) & 7) << 4;
  const uint32_t rowB = sZ + p * 128;
  const uint32_t kB = (p & 7) << 4;

  constexpr int NIDW = sizeof(CID) == 1 ? 4 : 8;
  auto fetch_ids = [&](int rr, int kk0, int jj0, uint32_t (&ids)[NIDW], uint32_t &lid_out) {
    const int kq = kk0 + pz, jq = jj0 + rw;
    lid_out = 0;
    if (jq < ny && kq < (rr ? rg.k1[1] : rg.k1[0])) {
      const uint4 *q4 = reinterpret_cast<const uint4 *>(cid + (((int64_t)kq * ny + jq) * NX + p * XW_M));
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
    xw_decode<XW_R>(t, rg, tiles_y, &r, &k0, &j0);
    fetch_ids(r, k0, j0, idw, lid);
  }
  uint32_t parC = 0, parZ = 0;
  bool stored = false;
  const int band_u = (ge_w - 1) >> 1;

  for (; t < n_tiles; t += n_groups) {
    const int j = j0 + rw;
    const int k1r = r ? rg.k1[1] : rg.k1[0];
    const bool line_ok = j < ny && k0 + pz < k1r;          // this lane's line exists and belongs to the launch
    // LPW = 2: the lower plane's line (half warp 0) exists whenever the warp has work
    const bool warp_ok = LPW == 1 ? line_ok : (j < ny && k0 < k1r);
    const bool more = t + n_groups < n_tiles;
    int rn = 0, k0n = 0, j0n = 0;
    if (more) xw_decode<XW_R>(t + n_groups, rg, tiles_y, &rn, &k0n, &j0n);

    // the private rows: plane k0-1 for a line of the lower plane, plane k0+2 for the upper plane
    // (below / above the grid: the neighbouring slab's halo plane, or zeros at a domain face)
    if (lane == 0 && warp_ok) {
      if (stored) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous d1 has left the rows
      const int n_ok = LPW == 1 ? 1 : (k0 + 1 < k1r ? 2 : 1);
      xw_expect_tx(bZ, n_ok * PART);
#pragma unroll
      for (int h = 0; h < LPW; ++h) {
        const int hz = LPW == 1 ? pz : h;
        if (h < n_ok) {
          const int zk = hz == 0 ? k0 - 1 : k0 + 2;
          const uint32_t dst = gb + 2 * XW_CBOX + (hz * XW_R + rw) * ROW + hx * PART;
          if (zk < 0 && has_halo_lo)
            xw_tma_load(dst, &tm.HloZ, bZ, j, 0, hx * 32);
          else if (zk >= nz && has_halo_hi)
            xw_tma_load(dst, &tm.HhiZ, bZ, j, 0, hx * 32);
          else
            xw_tma_load(dst, &tm.Z, bZ, j, zk, hx * 32);
        }
      }
    }
    xw_wait(bC, parC);
    parC ^= 1;

    double v[XW_M];
    double cB0 = 0.0, cBd = 0.0, cB15 = 0.0;
    bool slow = false;
    if (warp_ok) {
      slow = __any_sync(0xffffffffu, !xw_interior_one_class<CID>(idw));
      // centre values first: they are also the x neighbours
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double2 c = xw_lds(rowC + ((u << 4) ^ kC));
        v[2 * u] = c.x;
        v[2 * u + 1] = c.y;
      }
      // x neighbours across the chunk ends (closed outer faces: conductance 0, any finite value)
      double xl = xw_shfl_up<PW>(v[XW_M - 1], 1);
      double xr_end = xw_shfl_down<PW>(v[0], 1);
      if (WPL == 2) {
        // across the middle of the line: the neighbouring chunk's end value sits in the same box
        if (hx == 0 && lane == 31) {
          const int l = CR * 32 + rw + 1;
          xr_end = xw_lds1((pz == 0 ? sC0 : sC1) + l * 128 + ((l & 7) << 4));
        }
        if (hx == 1 && lane == 0) {
          const int l = CR * 31 + rw + 1;
          xl = xw_lds1((pz == 0 ? sC0 : sC1) + l * 128 + ((7 ^ (l & 7)) << 4) + 8);
        }
      }
      if (p == 0) xl = v[0];
      if (p == P - 1) xr_end = v[XW_M - 1];
      if (!slow) {
        // coefficient sets: cell 0, cells 1..14 (class of cell 8), cell 15.  y and z neighbours are read one
        // 16-byte unit ahead of the arithmetic (the loads keep their program order: volatile asm)
        const uint32_t q0 = coef_s + xw_cell_class<CID>(idw, 0) * (HS2_COEF_STRIDE * 8);
        const uint32_t qd = coef_s + xw_cell_class<CID>(idw, 8) * (HS2_COEF_STRIDE * 8);
        const uint32_t q15 = coef_s + xw_cell_class<CID>(idw, 15) * (HS2_COEF_STRIDE * 8);
        double2 ym = xw_lds(rowC - 128 + kYm), yp = xw_lds(rowC + 128 + kYp), za = xw_lds(rowA + kC);
        double cxm, cxp, cym, cyp, cA, cB;
        {
          const double2 a0 = xw_lds(q0), a1 = xw_lds(q0 + 16), a2 = xw_lds(q0 + 32);
          cxm = a0.x, cxp = a0.y, cym = a1.x, cyp = a1.y;
          cA = pz == 0 ? a2.y : a2.x;      // plane 0: z+ is in the patch, plane 1: z-
          cB = cB0 = pz == 0 ? a2.x : a2.y;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          double2 ymn = ym, ypn = yp, zan = za;
          if (u < 7) {
            ymn = xw_lds(rowC - 128 + (((u + 1) << 4) ^ kYm));
            ypn = xw_lds(rowC + 128 + (((u + 1) << 4) ^ kYp));
            zan = xw_lds(rowA + (((u + 1) << 4) ^ kC));
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int e = 2 * u + c;
            if (e == 1 || e == XW_M - 1) {
              const uint32_t qq = e == 1 ? qd : q15;
              const double2 a0 = xw_lds(qq), a1 = xw_lds(qq + 16), a2 = xw_lds(qq + 32);
              cxm = a0.x, cxp = a0.y, cym = a1.x, cyp = a1.y;
              cA = pz == 0 ? a2.y : a2.x;
              cB = pz == 0 ? a2.x : a2.y;
              if (e == 1)
                cBd = cB;
              else
                cB15 = cB;
            }
            const double tt = v[e];
            const double xr = e < XW_M - 1 ? v[e + 1] : xr_end;
            double rr = cxm * (xl - tt);
            rr = fma(cxp, xr - tt, rr);
            rr = fma(cym, (c ? ym.y : ym.x) - tt, rr);
            rr = fma(cyp, (c ? yp.y : yp.x) - tt, rr);
            rr = fma(cA, (c ? za.y : za.x) - tt, rr);
            rr = fma(-cB, tt, rr);             // the private row adds cB * z below
            xl = tt;
            v[e] = rr;
          }
          ym = ymn, yp = ypn, za = zan;
        }
      } else {
        // a chunk of the warp has a class change inside: one coefficient set per cell
        int last_id = -1;
        double cxm = 0, cxp = 0, cym = 0, cyp = 0, cA = 0, cB = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double2 ym = xw_lds(rowC - 128 + ((u << 4) ^ kYm));
          const double2 yp = xw_lds(rowC + 128 + ((u << 4) ^ kYp));
          const double2 za = xw_lds(rowA + ((u << 4) ^ kC));
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int e = 2 * u + c;
            const int idc = xw_cell_class<CID>(idw, e);
            if (idc != last_id) {
              const uint32_t qq = coef_s + idc * (HS2_COEF_STRIDE * 8);
              const double2 a0 = xw_lds(qq), a1 = xw_lds(qq + 16), a2 = xw_lds(qq + 32);
              cxm = a0.x, cxp = a0.y, cym = a1.x, cyp = a1.y;
              cA = pz == 0 ? a2.y : a2.x;
              cB = pz == 0 ? a2.x : a2.y;
              last_id = idc;
            }
            const double tt = v[e];
            const double xr = e < XW_M - 1 ? v[e + 1] : xr_end;
            double rr = cxm * (xl - tt);
            rr = fma(cxp, xr - tt, rr);
            rr = fma(cym, (c ? ym.y : ym.x) - tt, rr);
            rr = fma(cyp, (c ? yp.y : yp.x) - tt, rr);
            rr = fma(cA, (c ? za.y : za.x) - tt, rr);
            rr = fma(-cB, tt, rr);
            xl = tt;
            v[e] = rr;
          }
        }
      }
    }
    // this warp has read the shared boxes; the last of the group's warps to get here fetches the next patch
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      const int old = atomicAdd(cnt + g, 1);
      if (old == WPG - 1) {
        atomicExch(cnt + g, 0);
        if (more) issue_C(t + n_groups);
      }
    }
    if (more) fetch_ids(rn, k0n, j0n, idn, lidn);     // in flight during the solve

    if (warp_ok) {
      xw_wait(bZ, parZ);
      parZ ^= 1;
      // (LPW = 2, upper line absent: its private row was not loaded - it still holds finite values, the line's
      //  results are not stored)
      if (!slow) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double2 zb = xw_lds(rowB + ((u << 4) ^ kB));
          v[2 * u] = fma(u == 0 ? cB0 : cBd, zb.x, v[2 * u]);
          v[2 * u + 1] = fma(u == 7 ? cB15 : cBd, zb.y, v[2 * u + 1]);
        }
      } else {
        int last_id = -1;
        double cB = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double2 zb = xw_lds(rowB + ((u << 4) ^ kB));
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int e = 2 * u + c;
            const int idc = xw_cell_class<CID>(idw, e);
            if (idc != last_id) {
              const double2 a2 = xw_lds(coef_s + idc * (HS2_COEF_STRIDE * 8) + 32);
              cB = pz == 0 ? a2.x : a2.y;
              last_id = idc;
            }
            v[e] = fma(cB, c ? zb.y : zb.x, v[e]);
          }
        }
      }

      // ------------------------------------------------ partitioned solve along x, interfaces by shuffle
      // an absent line (LPW = 2) borrows the other half warp's unique-line id: same code path, results dropped
      const uint32_t lid_o = LPW == 1 ? lid : __shfl_xor_sync(0xffffffffu, lid, 16);     // (all lanes take part)
      const uint32_t lid_e = line_ok ? lid : lid_o;
      const bool uni_l = lid_e < (uint32_t)n_slots && s_code[lid_e] != 0;
      const bool uni = LPW == 1 ? uni_l : __all_sync(0xffffffffu, uni_l);     // warp-uniform
      // (y_first, y_last) of the chunks dl below / above this lane's: by shuffle inside the warp, through shared
      // memory across the middle of a two-warp line (a lane without that neighbour gets a finite value: its
      // interface coefficient is 0)
      double2 *sy = s_y + (g * 2 * XW_R + lw) * P;
      const int pair_id = 1 + g * 2 * XW_R + lw;
      auto nbr = [&](double yf_, double last_, int dl, double &yu, double &lu, double &yd, double &ld) {
        yu = xw_shfl_up<PW>(yf_, dl), lu = xw_shfl_up<PW>(last_, dl);
        yd = xw_shfl_down<PW>(yf_, dl), ld = xw_shfl_down<PW>(last_, dl);
        if (WPL == 2) {
          if (hx == 1 && lane < dl && p - dl >= 0) {
            const double2 t2 = sy[p - dl];
            yu = t2.x, lu = t2.y;
          }
          if (hx == 0 && lane + dl > 31 && p + dl < P) {
            const double2 t2 = sy[p + dl];
            yd = t2.x, ld = t2.y;
          }
        }
      };
      // alpha = E of the chunk before this lane's
      auto alpha_of = [&](double E) {
        double al = xw_shfl_up<PW>(E, 1);
        if (WPL == 2) {
          if (hx == 0 && lane == 31) s_e[g * 2 * XW_R + lw] = E;
          xw_pair_sync(pair_id);
          if (hx == 1 && lane == 0) al = s_e[g * 2 * XW_R + lw];
        }
        if (p == 0) al = 0.0;
        return al;
      };
      if (uni) {
        TabShared ts;
        ts.a = smem_u32(s_hdr + lid_e * XW_HDR);
        ts.pitch_b = XW_M * 8u;
        double last;
        const double yf = chunk_fwd<XW_M, true>(v, ts, XW_M, &last);
        const double2 *gr = s_ge + (int)lid_e * ge_w * P + p;            // [d][chunk]
        const double2 gc = gr[band_u * P];
        double e0 = gc.x * yf, e1 = gc.y * last, e2 = 0.0, e3 = 0.0;
        if (WPL == 2) {
          sy[p] = make_double2(yf, last);
          xw_pair_sync(pair_id);
        }
        for (int dl = 1; dl <= band_u; ++dl) {
          const double2 gu = gr[(band_u - dl) * P], gd = gr[(band_u + dl) * P];
          double yu, lu, yd, ld;
          nbr(yf, last, dl, yu, lu, yd, ld);
          e0 = fma(gu.x, yu, e0);
          e1 = fma(gu.y, lu, e1);
          e2 = fma(gd.x, yd, e2);
          e3 = fma(gd.y, ld, e3);
        }
        const double E = (e0 + e1) + (e2 + e3);
        const double alpha = alpha_of(E);
        chunk_bwd<XW_M, true>(v, ts, XW_M, alpha, E);
        if (p == 0) {
          // first chunk: its ghost neighbour is its own first value F: x = a - F b, F = a_0 / (1 + b_0)
          const uint32_t qb = ts.plane(HS2_T_PLANES);
          const double F = v[0] * TabShared::ld(ts.plane(HS2_T_PLANES + 1), 0);
#pragma unroll
          for (int q = 0; q < XW_M - 1; ++q) v[q] = fma(-F, TabShared::ld(qb, q), v[q]);
        }
      } else {
        TabGlobal tg;
        tg.b = tab + ((int64_t)lid_e * HS2_T_PLANES) * pitch + p * XW_M;
        tg.pitch = pitch;
        double last;
        const double yf = chunk_fwd<XW_M, true>(v, tg, XW_M, &last);
        const double2 *grow = reinterpret_cast<const double2 *>(GE + ((int64_t)lid_e * P + p) * (2 * P));
        const double2 gc = grow[p];
        double e0 = gc.x * yf, e1 = gc.y * last, e2 = 0.0, e3 = 0.0;
        if (WPL == 2) {
          sy[p] = make_double2(yf, last);
          xw_pair_sync(pair_id);
        }
        for (int dl = 1; dl <= band_g; ++dl) {
          const double2 gu = p - dl >= 0 ? grow[p - dl] : make_double2(0.0, 0.0);
          const double2 gd = p + dl < P ? grow[p + dl] : make_double2(0.0, 0.0);
          double yu, lu, yd, ld;
          nbr(yf, last, dl, yu, lu, yd, ld);
          e0 = fma(gu.x, yu, e0);
          e1 = fma(gu.y, lu, e1);
          e2 = fma(gd.x, yd, e2);
          e3 = fma(gd.y, ld, e3);
        }
        const double E = (e0 + e1) + (e2 + e3);
        const double alpha = alpha_of(E);
        chunk_bwd<XW_M, true>(v, tg, XW_M, alpha, E);
      }

      // ------------------------------------------------ d1 over the private row (same swizzle) -> bulk tensor store
      if (line_ok) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(rowB + ((u << 4) ^ kB)), "d"(v[2 * u]), "d"(v[2 * u + 1]) : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int h = 0; h < LPW; ++h) {
          const int hz = LPW == 1 ? pz : h;
          if (k0 + hz < k1r) xw_tma_store(&tm.O, gb + 2 * XW_CBOX + (hz * XW_R + rw) * ROW + hx * PART, j, k0 + hz, hx * 32);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      stored = true;
    }
    r = rn, k0 = k0n, j0 = j0n;
    lid = lidn;
#pragma unroll
    for (int q = 0; q < NIDW; ++q) idw[q] = idn[q];
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}


// ---------------------------------------------------------------------------------------------------------------
// Two lines per lane (nx = 512).  The profile of the kernel above (profiles/NOTES_r02.md) shows the L1 data pipe at
// 87 %: of 327 shared-memory wavefronts per line 160 are the five stencil streams and ~100 the broadcast loads of
// the chunk table (one wavefront per double whatever the width).  Here a warp owns row j of BOTH planes of the
// patch, lane p holds chunk p of the two lines:
//   * the in-patch z neighbour of one line is the centre value of the other - in registers (4 stencil streams
//     instead of 5);
//   * lines with the same unique-line id (all of them on the BASELINE grids) are solved together: every table
//     value and interface coefficient is loaded once for two lines, and the two recurrences interleave
//     (two independent DFMA chains per thread).
// A group is XW_R warps; shared memory per group as above.
template <int M>
__device__ __forceinline__ void chunk_fwd2(double (&a)[M], double (&b)[M], const TabShared &tb, double *yfa, double *la,
                                           double *yfb, double *lb) {
  {
    const uint32_t q = tb.plane(HS2_T_INV);
#pragma unroll
    for (int t = 0; t < M; ++t) {
      const double x = TabShared::ld(q, t);
      a[t] *= x;
      b[t] *= x;
    }
  }
  double pa = 0.0, pb = 0.0;
  {
    const uint32_t q = tb.plane(HS2_T_F);
#pragma unroll
    for (int t = 0; t < M; ++t) {
      const double f = TabShared::ld(q, t);
      pa = fma(-f, pa, a[t]);
      pb = fma(-f, pb, b[t]);
      a[t] = pa;
      b[t] = pb;
    }
  }
  *la = pa, *lb = pb;
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
  {
    const uint32_t q = tb.plane(HS2_T_C);
#pragma unroll
    for (int t = 0; t < M; t += 2) {
      const double c0 = TabShared::ld(q, t), c1 = TabShared::ld(q, t + 1);
      a0 = fma(c0, a[t], a0);
      a1 = fma(c1, a[t + 1], a1);
      b0 = fma(c0, b[t], b0);
      b1 = fma(c1, b[t + 1], b1);
    }
  }
  *yfa = a0 + a1, *yfb = b0 + b1;
}

template <int M>
__device__ __forceinline__ void chunk_bwd2(double (&a)[M], double (&b)[M], const TabShared &tb, double alpha_a, double Ea,
                                           double alpha_b, double Eb) {
  {
    const uint32_t q = tb.plane(HS2_T_S);
#pragma unroll
    for (int t = 0; t < M; ++t) {
      const double x = TabShared::ld(q, t);
      a[t] = fma(-alpha_a, x, a[t]);
      b[t] = fma(-alpha_b, x, b[t]);
    }
  }
  const uint32_t q = tb.plane(HS2_T_CP);
  double na = Ea, nb = Eb;
#pragma unroll
  for (int t = M - 1; t >= 0; --t) {
    if (t < M - 1) {
      const double x = TabShared::ld(q, t);
      na = fma(-x, na, a[t]);
      nb = fma(-x, nb, b[t]);
    }
    a[t] = na;
    b[t] = nb;
  }
}

template <typename CID, int XW_R, int XW_G>
__global__ void __maxnreg__(255)
sweep_xw2_kernel(const __grid_constant__ XwMaps tm, const CID *__restrict__ cid, const double *__restrict__ coef_g, int n_classes,
                 int has_halo_lo, int has_halo_hi, const uint32_t *__restrict__ line_id, const double *__restrict__ tab, int pitch,
                 const double *__restrict__ GE, int band_g, const double *__restrict__ xw_tab, const uint8_t *__restrict__ xw_code,
                 int n_slots, int ge_w, int nz, int ny, int tiles_y, int n_tiles, XwRanges rg) {
  constexpr int P = 32;
  constexpr int NX = P * XW_M;
  constexpr uint32_t ROW = P * 128;
  constexpr int WPG = XW_R;                                          // warps per group: one per row of the patch
  constexpr int CR = xw_box_rows(XW_R);
  constexpr uint32_t XW_CBOX = CR * ROW;
  constexpr uint32_t XW_GROUP = 2 * XW_CBOX + 2 * XW_R * ROW;
  extern __shared__ __align__(1024) unsigned char xsm[];
  double *s_hdr = reinterpret_cast<double *>(xsm + XW_G * XW_GROUP);               // [n_slots][XW_HDR]
  double2 *s_ge = reinterpret_cast<double2 *>(s_hdr + n_slots * XW_HDR);            // [n_slots][ge_w][P]
  double *cfs = reinterpret_cast<double *>(s_ge + n_slots * ge_w * P);               // [n_classes][8]
  uint64_t *barC = reinterpret_cast<uint64_t *>(cfs + n_classes * HS2_COEF_STRIDE);  // [XW_G]
  uint64_t *barZ = barC + XW_G;                                                      // [XW_G * WPG]
  int *cnt = reinterpret_cast<int *>(barZ + XW_G * WPG);                             // [XW_G]
  uint8_t *s_code = reinterpret_cast<uint8_t *>(cnt + XW_G);                         // [n_slots]

  const int tid = threadIdx.x;
  const int lane = tid & 31, wrp = tid >> 5;
  const int g = wrp / WPG, rw = wrp % WPG;
  const int p = lane;
  const int n_groups = gridDim.x * XW_G;

  const uint32_t gb = smem_u32(xsm) + g * XW_GROUP;
  const uint32_t sC0 = gb, sC1 = gb + XW_CBOX;
  const uint32_t sZ0 = gb + 2 * XW_CBOX + rw * ROW, sZ1 = sZ0 + XW_R * ROW;    // private rows of the two lines
  const uint32_t bC = smem_u32(barC + g), bZ = smem_u32(barZ + wrp);
  const uint32_t coef_s = smem_u32(cfs);

  if (tid == 0) {
    for (int q = 0; q < XW_G; ++q) {
      mbar_init(barC + q, 1);
      cnt[q] = 0;
    }
    for (int q = 0; q < XW_G * WPG; ++q) mbar_init(barZ + q, 1);
    fence_mbar_init();
  }
  {
    const int64_t stride = XW_HDR + ge_w * P * 2;
    for (int q = tid; q < n_slots * XW_HDR; q += blockDim.x) s_hdr[q] = xw_tab[(int64_t)(q / XW_HDR) * stride + q % XW_HDR];
    for (int q = tid; q < n_slots * ge_w * P; q += blockDim.x) {
      const int sl = q / (ge_w * P), e = q % (ge_w * P);
      s_ge[q] = reinterpret_cast<const double2 *>(xw_tab + (int64_t)sl * stride + XW_HDR)[e];
    }
    for (int q = tid; q < n_slots; q += blockDim.x) s_code[q] = xw_code[q];
    for (int q = tid; q < n_classes * HS2_COEF_STRIDE; q += blockDim.x) cfs[q] = coef_g[q];
  }
  __syncthreads();

  auto issue_C = [&](int t) {
    int r, k0, j0;
    xw_decode<XW_R>(t, rg, tiles_y, &r, &k0, &j0);
    xw_expect_tx(bC, 2 * XW_CBOX);
    xw_tma_load(sC0, &tm.C, bC, j0 - 1, k0);
    if (k0 + 1 >= nz && has_halo_hi)
      xw_tma_load(sC1, &tm.HhiC, bC, j0 - 1, 0);
    else
      xw_tma_load(sC1, &tm.C, bC, j0 - 1, k0 + 1);
  };

  int t = blockIdx.x * XW_G + g;
  if (t < n_tiles && rw == 0 && lane == 0) issue_C(t);

  const int lc = CR * p + rw + 1;
  const uint32_t row0 = sC0 + lc * 128, row1 = sC1 + lc * 128;
  const uint32_t kC = (lc & 7) << 4, kYm = ((lc - 1) & 7) << 4, kYp = ((lc + 1) & 7) << 4;
  const uint32_t rowB0 = sZ0 + p * 128, rowB1 = sZ1 + p * 128;
  const uint32_t kB = (p & 7) << 4;

  constexpr int NIDW = sizeof(CID) == 1 ? 4 : 8;
  auto fetch_ids = [&](int rr, int kq, int jq, uint32_t (&ids)[NIDW], uint32_t &lid_out) {
    lid_out = 0;
    if (jq < ny && kq < (rr ? rg.k1[1] : rg.k1[0])) {
      const uint4 *q4 = reinterpret_cast<const uint4 *>(cid + (((int64_t)kq * ny + jq) * NX + p * XW_M));
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
  uint32_t ia[NIDW], ib[NIDW], na[NIDW], nb[NIDW];       // class ids of this lane's chunk: lines A (plane k0), B (k0+1); next patch
  uint32_t lida = 0, lidb = 0, lidna = 0, lidnb = 0;
  if (t < n_tiles) {
    xw_decode<XW_R>(t, rg, tiles_y, &r, &k0, &j0);
    fetch_ids(r, k0, j0 + rw, ia, lida);
    fetch_ids(r, k0 + 1, j0 + rw, ib, lidb);
  }
  uint32_t parC = 0, parZ = 0;
  bool stored = false;
  const int band_u = (ge_w - 1) >> 1;

  // coefficient set of class c as seen from plane z of the patch: (x-, x+, y-, y+, in-patch z, out-of-patch z)
  struct Cf {
    double xm, xp, ym, yp, zi, zo;
  };
  auto load_cf = [&](int c, int z) {
    const uint32_t qq = coef_s + c * (HS2_COEF_STRIDE * 8);
    const double2 a0 = xw_lds(qq), a1 = xw_lds(qq + 16), a2 = xw_lds(qq + 32);
    Cf f;
    f.xm = a0.x, f.xp = a0.y, f.ym = a1.x, f.yp = a1.y;
    f.zi = z == 0 ? a2.y : a2.x;      // plane 0: z+ is in the patch, plane 1: z-
    f.zo = z == 0 ? a2.x : a2.y;
    return f;
  };

  for (; t < n_tiles; t += n_groups) {
    const int j = j0 + rw;
    const int k1r = r ? rg.k1[1] : rg.k1[0];
    const bool ok_a = j < ny && k0 < k1r;                 // line A exists whenever the warp has work
    const bool ok_b = j < ny && k0 + 1 < k1r;
    const bool more = t + n_groups < n_tiles;
    int rn = 0, k0n = 0, j0n = 0;
    if (more) xw_decode<XW_R>(t + n_groups, rg, tiles_y, &rn, &k0n, &j0n);

    // private rows: plane k0-1 for line A, plane k0+2 for line B (outside the grid: the neighbouring slab's halo
    // plane, or zeros at a domain face)
    if (lane == 0 && ok_a) {
      if (stored) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      xw_expect_tx(bZ, (ok_b ? 2 : 1) * ROW);
      if (k0 - 1 < 0 && has_halo_lo)
        xw_tma_load(sZ0, &tm.HloZ, bZ, j, 0);
      else
        xw_tma_load(sZ0, &tm.Z, bZ, j, k0 - 1);
      if (ok_b) {
        if (k0 + 2 >= nz && has_halo_hi)
          xw_tma_load(sZ1, &tm.HhiZ, bZ, j, 0);
        else
          xw_tma_load(sZ1, &tm.Z, bZ, j, k0 + 2);
      }
    }
    xw_wait(bC, parC);
    parC ^= 1;

    double va[XW_M], vb[XW_M];
    double oa0 = 0.0, oad = 0.0, oa15 = 0.0, ob0 = 0.0, obd = 0.0, ob15 = 0.0;     // out-of-patch z coefficients
    bool slow = false;
    if (ok_a) {
      slow = __any_sync(0xffffffffu, !xw_interior_one_class<CID>(ia) || !xw_interior_one_class<CID>(ib));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double2 ca = xw_lds(row0 + ((u << 4) ^ kC)), cb = xw_lds(row1 + ((u << 4) ^ kC));
        va[2 * u] = ca.x, va[2 * u + 1] = ca.y;
        vb[2 * u] = cb.x, vb[2 * u + 1] = cb.y;
      }
      double xla = xw_shfl_up<P>(va[XW_M - 1], 1), xra_end = xw_shfl_down<P>(va[0], 1);
      double xlb = xw_shfl_up<P>(vb[XW_M - 1], 1), xrb_end = xw_shfl_down<P>(vb[0], 1);
      if (p == 0) xla = va[0], xlb = vb[0];
      if (p == P - 1) xra_end = va[XW_M - 1], xrb_end = vb[XW_M - 1];
      if (!slow) {
        const int ca0 = xw_cell_class<CID>(ia, 0), cad = xw_cell_class<CID>(ia, 8), ca15 = xw_cell_class<CID>(ia, 15);
        const int cb0 = xw_cell_class<CID>(ib, 0), cbd = xw_cell_class<CID>(ib, 8), cb15 = xw_cell_class<CID>(ib, 15);
        Cf fa = load_cf(ca0, 0), fb = load_cf(cb0, 1);
        oa0 = fa.zo, ob0 = fb.zo;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double2 yma = xw_lds(row0 - 128 + ((u << 4) ^ kYm)), ypa = xw_lds(row0 + 128 + ((u << 4) ^ kYp));
          const double2 ymb = xw_lds(row1 - 128 + ((u << 4) ^ kYm)), ypb = xw_lds(row1 + 128 + ((u << 4) ^ kYp));
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int e = 2 * u + c;
            if (e == 1) {
              fa = load_cf(cad, 0), fb = load_cf(cbd, 1);
              oad = fa.zo, obd = fb.zo;
            }
            if (e == XW_M - 1) {
              fa = load_cf(ca15, 0), fb = load_cf(cb15, 1);
              oa15 = fa.zo, ob15 = fb.zo;
            }
            const double ta = va[e], tb = vb[e];
            const double xra = e < XW_M - 1 ? va[e + 1] : xra_end;
            const double xrb = e < XW_M - 1 ? vb[e + 1] : xrb_end;
            double ra = fa.xm * (xla - ta), rb = fb.xm * (xlb - tb);
            ra = fma(fa.xp, xra - ta, ra), rb = fma(fb.xp, xrb - tb, rb);
            ra = fma(fa.ym, (c ? yma.y : yma.x) - ta, ra), rb = fma(fb.ym, (c ? ymb.y : ymb.x) - tb, rb);
            ra = fma(fa.yp, (c ? ypa.y : ypa.x) - ta, ra), rb = fma(fb.yp, (c ? ypb.y : ypb.x) - tb, rb);
            ra = fma(fa.zi, tb - ta, ra), rb = fma(fb.zi, ta - tb, rb);
            ra = fma(-fa.zo, ta, ra), rb = fma(-fb.zo, tb, rb);       // the private rows add zo * z below
            xla = ta, xlb = tb;
            va[e] = ra, vb[e] = rb;
          }
        }
      } else {
        // a chunk of the warp has a class change inside: one coefficient set per cell
        int last_a = -1, last_b = -1;
        Cf fa = {0, 0, 0, 0, 0, 0}, fb = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double2 yma = xw_lds(row0 - 128 + ((u << 4) ^ kYm)), ypa = xw_lds(row0 + 128 + ((u << 4) ^ kYp));
          const double2 ymb = xw_lds(row1 - 128 + ((u << 4) ^ kYm)), ypb = xw_lds(row1 + 128 + ((u << 4) ^ kYp));
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int e = 2 * u + c;
            const int ida = xw_cell_class<CID>(ia, e), idb = xw_cell_class<CID>(ib, e);
            if (ida != last_a) fa = load_cf(ida, 0), last_a = ida;
            if (idb != last_b) fb = load_cf(idb, 1), last_b = idb;
            const double ta = va[e], tb = vb[e];
            const double xra = e < XW_M - 1 ? va[e + 1] : xra_end;
            const double xrb = e < XW_M - 1 ? vb[e + 1] : xrb_end;
            double ra = fa.xm * (xla - ta), rb = fb.xm * (xlb - tb);
            ra = fma(fa.xp, xra - ta, ra), rb = fma(fb.xp, xrb - tb, rb);
            ra = fma(fa.ym, (c ? yma.y : yma.x) - ta, ra), rb = fma(fb.ym, (c ? ymb.y : ymb.x) - tb, rb);
            ra = fma(fa.yp, (c ? ypa.y : ypa.x) - ta, ra), rb = fma(fb.yp, (c ? ypb.y : ypb.x) - tb, rb);
            ra = fma(fa.zi, tb - ta, ra), rb = fma(fb.zi, ta - tb, rb);
            ra = fma(-fa.zo, ta, ra), rb = fma(-fb.zo, tb, rb);
            xla = ta, xlb = tb;
            va[e] = ra, vb[e] = rb;
          }
        }
      }
    }
    // this warp has read the shared boxes; the last of the group's warps to get here fetches the next patch
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      const int old = atomicAdd(cnt + g, 1);
      if (old == WPG - 1) {
        atomicExch(cnt + g, 0);
        if (more) issue_C(t + n_groups);
      }
    }
    if (more) {      // in flight during the solve
      fetch_ids(rn, k0n, j0n + rw, na, lidna);
      fetch_ids(rn, k0n + 1, j0n + rw, nb, lidnb);
    }

    if (ok_a) {
      xw_wait(bZ, parZ);
      parZ ^= 1;
      // (line B absent: its private row was not loaded - it still holds finite values, the line's results are not stored)
      if (!slow) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double2 za = xw_lds(rowB0 + ((u << 4) ^ kB)), zb = xw_lds(rowB1 + ((u << 4) ^ kB));
          va[2 * u] = fma(u == 0 ? oa0 : oad, za.x, va[2 * u]);
          va[2 * u + 1] = fma(u == 7 ? oa15 : oad, za.y, va[2 * u + 1]);
          vb[2 * u] = fma(u == 0 ? ob0 : obd, zb.x, vb[2 * u]);
          vb[2 * u + 1] = fma(u == 7 ? ob15 : obd, zb.y, vb[2 * u + 1]);
        }
      } else {
        int last_a = -1, last_b = -1;
        double oa = 0, ob = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double2 za = xw_lds(rowB0 + ((u << 4) ^ kB)), zb = xw_lds(rowB1 + ((u << 4) ^ kB));
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int e = 2 * u + c;
            const int ida = xw_cell_class<CID>(ia, e), idb = xw_cell_class<CID>(ib, e);
            if (ida != last_a) oa = xw_lds(coef_s + ida * (HS2_COEF_STRIDE * 8) + 32).x, last_a = ida;
            if (idb != last_b) ob = xw_lds(coef_s + idb * (HS2_COEF_STRIDE * 8) + 32).y, last_b = idb;
            va[e] = fma(oa, c ? za.y : za.x, va[e]);
            vb[e] = fma(ob, c ? zb.y : zb.x, vb[e]);
          }
        }
      }

      // ------------------------------------------------ partitioned solve along x, interfaces by shuffle
      const uint32_t la_id = lida, lb_id = ok_b ? lidb : lida;       // an absent line B borrows A's tables (results dropped)
      const bool uni_a = la_id < (uint32_t)n_slots && s_code[la_id] != 0;
      const bool uni_b = lb_id < (uint32_t)n_slots && s_code[lb_id] != 0;
      // one line with whatever tables it has
      auto solve_one = [&](double (&v)[XW_M], uint32_t lid_e, bool uni) {
        if (uni) {
          TabShared ts;
          ts.a = smem_u32(s_hdr + lid_e * XW_HDR);
          ts.pitch_b = XW_M * 8u;
          double last;
          const double yf = chunk_fwd<XW_M, true>(v, ts, XW_M, &last);
          const double2 *gr = s_ge + (int)lid_e * ge_w * P + p;
          const double2 gc = gr[band_u * P];
          double e0 = gc.x * yf, e1 = gc.y * last, e2 = 0.0, e3 = 0.0;
          for (int dl = 1; dl <= band_u; ++dl) {
            const double2 gu = gr[(band_u - dl) * P], gd = gr[(band_u + dl) * P];
            e0 = fma(gu.x, xw_shfl_up<P>(yf, dl), e0);
            e1 = fma(gu.y, xw_shfl_up<P>(last, dl), e1);
            e2 = fma(gd.x, xw_shfl_down<P>(yf, dl), e2);
            e3 = fma(gd.y, xw_shfl_down<P>(last, dl), e3);
          }
          const double E = (e0 + e1) + (e2 + e3);
          double alpha = xw_shfl_up<P>(E, 1);
          if (p == 0) alpha = 0.0;
          chunk_bwd<XW_M, true>(v, ts, XW_M, alpha, E);
          if (p == 0) {
            const uint32_t qb = ts.plane(HS2_T_PLANES);
            const double F = v[0] * TabShared::ld(ts.plane(HS2_T_PLANES + 1), 0);
#pragma unroll
            for (int q = 0; q < XW_M - 1; ++q) v[q] = fma(-F, TabShared::ld(qb, q), v[q]);
          }
        } else {
          TabGlobal tg;
          tg.b = tab + ((int64_t)lid_e * HS2_T_PLANES) * pitch + p * XW_M;
          tg.pitch = pitch;
          double last;
          const double yf = chunk_fwd<XW_M, true>(v, tg, XW_M, &last);
          const double2 *grow = reinterpret_cast<const double2 *>(GE + ((int64_t)lid_e * P + p) * (2 * P));
          const double2 gc = grow[p];
          double e0 = gc.x * yf, e1 = gc.y * last, e2 = 0.0, e3 = 0.0;
          for (int dl = 1; dl <= band_g; ++dl) {
            const double2 gu = p - dl >= 0 ? grow[p - dl] : make_double2(0.0, 0.0);
            const double2 gd = p + dl < P ? grow[p + dl] : make_double2(0.0, 0.0);
            e0 = fma(gu.x, xw_shfl_up<P>(yf, dl), e0);
            e1 = fma(gu.y, xw_shfl_up<P>(last, dl), e1);
            e2 = fma(gd.x, xw_shfl_down<P>(yf, dl), e2);
            e3 = fma(gd.y, xw_shfl_down<P>(last, dl), e3);
          }
          const double E = (e0 + e1) + (e2 + e3);
          double alpha = xw_shfl_up<P>(E, 1);
          if (p == 0) alpha = 0.0;
          chunk_bwd<XW_M, true>(v, tg, XW_M, alpha, E);
        }
      };
      if (uni_a && la_id == lb_id) {
        // both lines on one ghost-uniform table: every table value is loaded once for the two of them
        TabShared ts;
        ts.a = smem_u32(s_hdr + la_id * XW_HDR);
        ts.pitch_b = XW_M * 8u;
        double yfa, lsa, yfb, lsb;
        chunk_fwd2<XW_M>(va, vb, ts, &yfa, &lsa, &yfb, &lsb);
        const double2 *gr = s_ge + (int)la_id * ge_w * P + p;
        const double2 gc = gr[band_u * P];
        double a0 = gc.x * yfa, a1 = gc.y * lsa, a2 = 0.0, a3 = 0.0;
        double b0 = gc.x * yfb, b1 = gc.y * lsb, b2 = 0.0, b3 = 0.0;
        for (int dl = 1; dl <= band_u; ++dl) {
          const double2 gu = gr[(band_u - dl) * P], gd = gr[(band_u + dl) * P];
          a0 = fma(gu.x, xw_shfl_up<P>(yfa, dl), a0);
          a1 = fma(gu.y, xw_shfl_up<P>(lsa, dl), a1);
          a2 = fma(gd.x, xw_shfl_down<P>(yfa, dl), a2);
          a3 = fma(gd.y, xw_shfl_down<P>(lsa, dl), a3);
          b0 = fma(gu.x, xw_shfl_up<P>(yfb, dl), b0);
          b1 = fma(gu.y, xw_shfl_up<P>(lsb, dl), b1);
          b2 = fma(gd.x, xw_shfl_down<P>(yfb, dl), b2);
          b3 = fma(gd.y, xw_shfl_down<P>(lsb, dl), b3);
        }
        const double Ea = (a0 + a1) + (a2 + a3), Eb = (b0 + b1) + (b2 + b3);
        double ala = xw_shfl_up<P>(Ea, 1), alb = xw_shfl_up<P>(Eb, 1);
        if (p == 0) ala = 0.0, alb = 0.0;
        chunk_bwd2<XW_M>(va, vb, ts, ala, Ea, alb, Eb);
        if (p == 0) {
          const uint32_t qb = ts.plane(HS2_T_PLANES);
          const double f1 = TabShared::ld(ts.plane(HS2_T_PLANES + 1), 0);
          const double Fa = va[0] * f1, Fb = vb[0] * f1;
#pragma unroll
          for (int q = 0; q < XW_M - 1; ++q) {
            const double x = TabShared::ld(qb, q);
            va[q] = fma(-Fa, x, va[q]);
            vb[q] = fma(-Fb, x, vb[q]);
          }
        }
      } else {
        solve_one(va, la_id, uni_a);
        solve_one(vb, lb_id, uni_b);
      }

      // ------------------------------------------------ d1 over the private rows (same swizzle) -> bulk tensor stores
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(rowB0 + ((u << 4) ^ kB)), "d"(va[2 * u]), "d"(va[2 * u + 1]) : "memory");
      if (ok_b) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(rowB1 + ((u << 4) ^ kB)), "d"(vb[2 * u]), "d"(vb[2 * u + 1]) : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        xw_tma_store(&tm.O, sZ0, j, k0);
        if (ok_b) xw_tma_store(&tm.O, sZ1, j, k0 + 1);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      stored = true;
    }
    r = rn, k0 = k0n, j0 = j0n;
    lida = lidna, lidb = lidnb;
#pragma unroll
    for (int q = 0; q < NIDW; ++q) ia[q] = na[q], ib[q] = nb[q];
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

bool xw_encode4(CUtensorMap *m, const void *base, int nx, int ny, int nzz, int rows, int segs = 0) {
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
  cuuint32_t box[4] = {16, (cuuint32_t)rows, 1, (cuuint32_t)(segs ? segs : S)};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void *>(base), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename CID, int XW_R, int XW_G, int LPW, int WPL = 1>
int launch_xw(hs2_plan *p, const double *T, double *W, const double *halo_lo, const double *halo_hi, int part,
              cudaStream_t st, bool *done) {
  constexpr int P = 32 * WPL / LPW, WPG = 2 * XW_R * WPL / LPW, SEG = P / WPL;
  constexpr uint32_t XW_GROUP = (2 * xw_box_rows(XW_R) + 2 * XW_R) * P * 128;
  *done = false;
  const hs2_plan_desc &d = p->d;
  const hs2_axis_tables &ax = d.axis[0];
  const int nx = (int)d.nx, ny = (int)d.ny, nz = (int)d.nz;
  XwRanges rg;
  if (part == 0) {
    rg.n = 1, rg.k0[0] = 0, rg.k1[0] = nz, rg.k0[1] = rg.k1[1] = 0;
  } else if (part == HS2_X_INTERIOR) {
    rg.n = 1, rg.k0[0] = 1, rg.k1[0] = nz - 1, rg.k0[1] = rg.k1[1] = 0;
  } else {
    rg.n = 2, rg.k0[0] = 0, rg.k1[0] = 1, rg.k0[1] = nz - 1, rg.k1[1] = nz;
  }
  const int tiles_y = (ny + XW_R - 1) / XW_R;
  int64_t n_tiles = 0;
  for (int r = 0; r < rg.n; ++r) n_tiles += (int64_t)((rg.k1[r] - rg.k0[r] + 1) / 2) * tiles_y;
  if (n_tiles <= 0) {
    *done = true;
    return HS2_OK;
  }
  if (n_tiles >= ((int64_t)1 << 31)) return HS2_OK;
  XwMaps tm;
  memset(&tm, 0, sizeof(tm));
  if (!xw_encode4(&tm.C, T, nx, ny, nz, xw_box_rows(XW_R)) || !xw_encode4(&tm.Z, T, nx, ny, nz, 1, SEG) ||
      !xw_encode4(&tm.O, W, nx, ny, nz, 1, SEG))
    return HS2_OK;
  if (halo_lo && !xw_encode4(&tm.HloZ, halo_lo, nx, ny, 1, 1, SEG)) return HS2_OK;
  if (halo_hi && (!xw_encode4(&tm.HhiC, halo_hi, nx, ny, 1, xw_box_rows(XW_R)) || !xw_encode4(&tm.HhiZ, halo_hi, nx, ny, 1, 1, SEG)))
    return HS2_OK;
  const int ge_w = 2 * ax.xw_band + 1;
  const size_t per_slot = (size_t)XW_HDR * 8 + (size_t)ge_w * P * 16 + 1;
  const size_t fixed = (size_t)XW_G * XW_GROUP + (size_t)d.n_classes * HS2_COEF_STRIDE * 8 +
                       (size_t)(XW_G + XW_G * WPG) * 8 + (WPL == 2 ? (size_t)XW_G * 2 * XW_R * (P * 16 + 8) : 0) + XW_G * 4 + 64;
  if (fixed + 1024 > (size_t)p->max_smem_optin) return HS2_OK;
  int n_slots = (int)(((size_t)p->max_smem_optin - 1024 - fixed) / per_slot);
  if (n_slots > ax.n_unique) n_slots = ax.n_unique;
  if (n_slots > 64) n_slots = 64;
  const size_t smem = fixed + (size_t)n_slots * per_slot;
  auto kern = sweep_xw_kernel<CID, XW_R, XW_G, LPW, WPL>;
  HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = p->sm_count;
  const int64_t need = (n_tiles + XW_G - 1) / XW_G;
  if (blocks > need) blocks = need;
  kern<<<(unsigned)blocks, XW_G * WPG * 32, smem, st>>>(tm, (const CID *)d.d_class_id, d.d_class_coef, d.n_classes,
                                                        halo_lo ? 1 : 0, halo_hi ? 1 : 0, ax.d_line_id, ax.d_tab, ax.pitch, ax.d_GE,
                                                        ax.band, ax.d_xw_tab, ax.d_xw_code, n_slots, ge_w, nz, ny, tiles_y,
                                                        (int)n_tiles, rg);
  HS2_CUDA_CHECK(cudaGetLastError());
  p->last_kernel[0] = HS2_K_X_WARP;
  *done = true;
  return HS2_OK;
}

template <typename CID, int XW_R, int XW_G>
int launch_xw2(hs2_plan *p, const double *T, double *W, const double *halo_lo, const double *halo_hi, int part,
              cudaStream_t st, bool *done) {
  constexpr int P = 32, WPG = XW_R, SEG = P, WPL = 1;
  constexpr uint32_t XW_GROUP = (2 * xw_box_rows(XW_R) + 2 * XW_R) * P * 128;
  *done = false;
  const hs2_plan_desc &d = p->d;
  const hs2_axis_tables &ax = d.axis[0];
  const int nx = (int)d.nx, ny = (int)d.ny, nz = (int)d.nz;
  XwRanges rg;
  if (part == 0) {
    rg.n = 1, rg.k0[0] = 0, rg.k1[0] = nz, rg.k0[1] = rg.k1[1] = 0;
  } else if (part == HS2_X_INTERIOR) {
    rg.n = 1, rg.k0[0] = 1, rg.k1[0] = nz - 1, rg.k0[1] = rg.k1[1] = 0;
  } else {
    rg.n = 2, rg.k0[0] = 0, rg.k1[0] = 1, rg.k0[1] = nz - 1, rg.k1[1] = nz;
  }
  const int tiles_y = (ny + XW_R - 1) / XW_R;
  int64_t n_tiles = 0;
  for (int r = 0; r < rg.n; ++r) n_tiles += (int64_t)((rg.k1[r] - rg.k0[r] + 1) / 2) * tiles_y;
  if (n_tiles <= 0) {
    *done = true;
    return HS2_OK;
  }
  if (n_tiles >= ((int64_t)1 << 31)) return HS2_OK;
  XwMaps tm;
  memset(&tm, 0, sizeof(tm));
  if (!xw_encode4(&tm.C, T, nx, ny, nz, xw_box_rows(XW_R)) || !xw_encode4(&tm.Z, T, nx, ny, nz, 1, SEG) ||
      !xw_encode4(&tm.O, W, nx, ny, nz, 1, SEG))
    return HS2_OK;
  if (halo_lo && !xw_encode4(&tm.HloZ, halo_lo, nx, ny, 1, 1, SEG)) return HS2_OK;
  if (halo_hi && (!xw_encode4(&tm.HhiC, halo_hi, nx, ny, 1, xw_box_rows(XW_R)) || !xw_encode4(&tm.HhiZ, halo_hi, nx, ny, 1, 1, SEG)))
    return HS2_OK;
  const int ge_w = 2 * ax.xw_band + 1;
  const size_t per_slot = (size_t)XW_HDR * 8 + (size_t)ge_w * P * 16 + 1;
  const size_t fixed = (size_t)XW_G * XW_GROUP + (size_t)d.n_classes * HS2_COEF_STRIDE * 8 +
                       (size_t)(XW_G + XW_G * WPG) * 8 + (WPL == 2 ? (size_t)XW_G * 2 * XW_R * (P * 16 + 8) : 0) + XW_G * 4 + 64;
  if (fixed + 1024 > (size_t)p->max_smem_optin) return HS2_OK;
  int n_slots = (int)(((size_t)p->max_smem_optin - 1024 - fixed) / per_slot);
  if (n_slots > ax.n_unique) n_slots = ax.n_unique;
  if (n_slots > 64) n_slots = 64;
  const size_t smem = fixed + (size_t)n_slots * per_slot;
  auto kern = sweep_xw2_kernel<CID, XW_R, XW_G>;
  HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = p->sm_count;
  const int64_t need = (n_tiles + XW_G - 1) / XW_G;
  if (blocks > need) blocks = need;
  kern<<<(unsigned)blocks, XW_G * WPG * 32, smem, st>>>(tm, (const CID *)d.d_class_id, d.d_class_coef, d.n_classes,
                                                        halo_lo ? 1 : 0, halo_hi ? 1 : 0, ax.d_line_id, ax.d_tab, ax.pitch, ax.d_GE,
                                                        ax.band, ax.d_xw_tab, ax.d_xw_code, n_slots, ge_w, nz, ny, tiles_y,
                                                        (int)n_tiles, rg);
  HS2_CUDA_CHECK(cudaGetLastError());
  p->last_kernel[0] = HS2_K_X_WARP;
  *done = true;
  return HS2_OK;
}

template <typename CID>
int launch_xw_v(hs2_plan *p, const double *T, double *W, const double *halo_lo, const double *halo_hi, int part,
                cudaStream_t st, bool *done) {
  // HS2_XW_SHAPE: 33 = patches of 2 x 3 rows, three groups per block (18 warps, 96 registers),
  //               52 = patches of 2 x 5 rows, two groups (20 warps, 96 registers),
  //               42 = patches of 2 x 4 rows (7-row boxes), two groups (16 warps, 128 registers)
  // (a scheduler's quarter of the register file holds 4 warps of 128 or 5 of 96 registers)
  const int shape = getenv("HS2_XW_SHAPE") ? atoi(getenv("HS2_XW_SHAPE")) : 24;
  if (p->d.nx == 256) return launch_xw<CID, 4, 4, 2>(p, T, W, halo_lo, halo_hi, part, st, done);   // 2 lines per warp
  if (p->d.nx == 1024) {
    // 2 warps per line, rows of 8 KB: patches of 2 x 1 rows (3-row boxes), two groups per block (8 warps);
    // HS2_XW_SHAPE=13: three groups (12 warps) - measured slower (0.79 vs 0.76 ms on 128 x 1024 x 1024)
    if (shape == 13) return launch_xw<CID, 1, 3, 1, 2>(p, T, W, halo_lo, halo_hi, part, st, done);
    return launch_xw<CID, 1, 2, 1, 2>(p, T, W, halo_lo, halo_hi, part, st, done);
  }
  // 24 (default): two lines per lane (sweep_xw2_kernel), patches of 2 x 4 rows, two groups per block (8 warps of
  // 255 registers) - 0.524 ms at 512^3 against 0.561 for 42; 42 / 33 / 52: one line per lane
  if (shape == 24) return launch_xw2<CID, 4, 2>(p, T, W, halo_lo, halo_hi, part, st, done);
  if (shape == 52) return launch_xw<CID, 5, 2, 1>(p, T, W, halo_lo, halo_hi, part, st, done);
  if (shape == 33) return launch_xw<CID, 3, 3, 1>(p, T, W, halo_lo, halo_hi, part, st, done);
  return launch_xw<CID, 4, 2, 1>(p, T, W, halo_lo, halo_hi, part, st, done);
}

}  // namespace

bool hs2_tile_xw_supported(const hs2_plan *p) {
  const hs2_plan_desc &d = p->d;
  const hs2_axis_tables &ax = d.axis[0];
  if (d.flags & (HS2_FLAG_FORCE_FALLBACK | HS2_FLAG_X_FOLD | HS2_FLAG_X_PATCH)) return false;
  if (ax.chunk != XW_M || !ax.d_tab || !ax.d_GE || ax.pitch <= 0 || !ax.d_xw_tab || !ax.d_xw_code || ax.xw_band < 0) return false;
  if ((d.nx != 1024 && d.nx != 512 && d.nx != 256) || ax.n_chunks != d.nx / XW_M || d.n_classes > 64) return false;
  if (d.ny >= ((int64_t)1 << 30) || d.nz >= ((int64_t)1 << 30)) return false;
  if ((reinterpret_cast<uintptr_t>(d.d_class_id) & 15)) return false;
  return true;
}

int hs2_tile_sweep_xw(hs2_plan *p, const double *T, double *W, const double *halo_lo, const double *halo_hi, int part,
                      cudaStream_t st, bool *done) {
  *done = false;
  if (p->d.class_id_bytes == 1) return launch_xw_v<uint8_t>(p, T, W, halo_lo, halo_hi, part, st, done);
  return launch_xw_v<uint16_t>(p, T, W, halo_lo, halo_hi, part, st, done);
}

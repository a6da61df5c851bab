// kernels_strided.cu - y- and z-sweeps: batched tridiagonal solves along a
// strided axis in ONE pass over HBM (read 8 B + write 8 B per cell; the z-sweep
// also reads T_in and writes T_out = T_in + increment).
//
// Layout: a block owns a tile of W adjacent lines (W contiguous x positions,
// so every row of the tile is one W*8-byte segment) over the full line length
// L.  The tile is cut into P chunks of M rows; thread (w, p) keeps chunk p of
// line w in registers (M doubles).  A warp is 32/W chunks of the same W lines,
// so the per-row factors - which depend on the line's *class* only - are
// (nearly) warp-uniform loads served from L1.
//
// Algorithm per chunk (tables and derivation: heatsim2_b200/plan.py,
// chunk_factors): local forward elimination from a zero entry value; the first
// and last rows of the local solution go to shared memory; every thread applies
// one row of the precomputed inverse of the 2P x 2P interface system to get the
// true value of its last row, E_p; alpha = E_{p-1} comes from the neighbour
// through shared memory; back substitution from the known last row.
// The phases are written so that all loads of a phase are independent of the
// only serial part (one DFMA per row per direction).
// This replaces heatsim2/tridiag.pyx:46-69 (one serial chain over the whole
// grid) and the transposes of alternatingdirection_c_pyx.pyx:397,412.
#include "chunk_core.cuh"
#include "tma_util.cuh"
#include <stdlib.h>

namespace {

template <int M, int W, bool FINAL>
__global__ void __launch_bounds__(M >= 32 ? 256 : 512, 2)
strided_sweep(double *__restrict__ data, const double *__restrict__ Tin, double *__restrict__ Tout,
              const uint32_t *__restrict__ line_id, const double *__restrict__ tab, const double *__restrict__ GE,
              int L, int pitch, int P, int band,
              int64_t stride,          // elements between consecutive rows of a line
              int tiles_per_group,     // tiles per contiguous group of lines
              int lines_per_group,     // lines in a group (y: nx, z: ny*nx)
              int64_t group_stride)    // elements between groups (y: ny*nx, z: unused)
{
  extern __shared__ double sm[];         // Y[2P][W] then E[P][W]
  double *Y = sm;
  double *Es = sm + 2 * P * W;
  const int w = threadIdx.x;             // line within tile
  const int p = threadIdx.y;             // chunk
  const int group = blockIdx.x / tiles_per_group;
  const int col = (blockIdx.x % tiles_per_group) * W + w;
  const bool live = col < lines_per_group;
  const int64_t line = (int64_t)group * lines_per_group + (live ? col : 0);
  const int64_t base = (int64_t)group * group_stride + (live ? col : 0);
  const int r0 = p * M;
  const int rows = min(M, L - r0);       // >= 1 for every launched chunk
  const bool full = rows == M;
  const uint32_t lid = line_id[line];
  const double *tb = tab + ((int64_t)lid * HS2_T_PLANES) * pitch + r0;
  const double *ge = GE + ((int64_t)lid * P + p) * (2 * P);
  const int64_t off = base + (int64_t)r0 * stride;

  double v[M];
  double yf, last;
  if (full) {
    const double *src = data + off;
#pragma unroll
    for (int t = 0; t < M; ++t) {
      v[t] = live ? *src : 0.0;
      src += stride;
    }
    yf = chunk_forward_full<M>(v, tb, pitch);
    last = v[M - 1];
  } else {
#pragma unroll
    for (int t = 0; t < M; ++t) v[t] = (live && t < rows) ? data[off + (int64_t)t * stride] : 0.0;
    yf = chunk_forward_short<M>(v, tb, pitch, rows, &last);
  }
  Y[(2 * p) * W + w] = yf;
  Y[(2 * p + 1) * W + w] = last;
  __syncthreads();
  const double E = chunk_interface(ge, Y, P, W, w, p, band);
  Es[p * W + w] = E;
  __syncthreads();
  const double alpha = p > 0 ? Es[(p - 1) * W + w] : 0.0;
  if (full) {
    chunk_backward_full<M>(v, tb, pitch, alpha, E);
    if (live) {
      if (FINAL) {
        const double *ti = Tin + off;
        double *to = Tout + off;
#pragma unroll
        for (int t = 0; t < M; ++t) {
          *to = *ti + v[t];
          ti += stride;
          to += stride;
        }
      } else {
        double *dst = data + off;
#pragma unroll
        for (int t = 0; t < M; ++t) {
          *dst = v[t];
          dst += stride;
        }
      }
    }
  } else {
    chunk_backward_short<M>(v, tb, pitch, rows, alpha, E);
    if (live) {
#pragma unroll
      for (int t = 0; t < M; ++t) {
        if (t < rows) {
          const int64_t a = off + (int64_t)t * stride;
          if (FINAL)
            Tout[a] = Tin[a] + v[t];
          else
            data[a] = v[t];
        }
      }
    }
  }
}

// Persistent variant: a block walks tiles blockIdx.x, +gridDim.x, ...; the
// next tile is fetched into shared memory by the TMA engine
// (cp.async.bulk.tensor) while the current one is being solved from registers,
// so HBM reads never wait for arithmetic.  The tile lands as dense rows
// [L_pad][W]; thread (w, p) copies its chunk to registers, after which the
// buffer is handed back to the copy engine.  Results are stored straight from
// registers.  Tensor map: rank 3, dims (n0, n1, n2) = (nx, ny, nz) for the
// y-sweep and (ny*nx, nz, 1) for the z-sweep, box (W, BR, 1).
// UT: warps whose chunks all carry the axis' most common table read the factors as constant operands
// (chunk_core.cuh).  Measured on B200 at 512^3 (scripts/ab_sweeps.py, one run): z sweep 0.641 -> 0.613 ms,
// y sweep 0.377 -> 0.401 ms (it is HBM-bound already and the uniform loads add latency) - so UT is
// instantiated for the z sweep only.
// The 16-byte cp.async pieces of one thread for one tile [L][W] (row r of the tile <- src + r * stride): thread tid
// copies piece tid % (W/2) of rows tid / (W/2), + 2 nthr / W, ...  one() issues the next piece, rest() what is left
// and closes the group; a default-constructed pacer does nothing.
struct CpAsyncPacer {
  const double *src = nullptr;
  uint32_t dst = 0, dstep = 0;
  int64_t sstep = 0;
  int left = 0, ok = 0;
  bool armed = false;
  __device__ __forceinline__ void setup(const double *src0, const double *tile, int tid, int W, int nthr, int L, int64_t stride,
                                        int lines_left) {
    const int c = (tid % (W / 2)) * 2;
    const int rstep = nthr / (W / 2);
    const int r = tid / (W / 2);
    ok = c < lines_left ? 16 : 0;
    src = src0 + (int64_t)r * stride + (ok ? c : 0);
    dst = smem_u32(tile + (size_t)r * W + c);
    sstep = (int64_t)rstep * stride;
    dstep = (uint32_t)rstep * (uint32_t)(W * 8);
    left = r < L ? (L - r + rstep - 1) / rstep : 0;
    armed = true;
  }
  __device__ __forceinline__ void one() {
    if (left > 0) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok) : "memory");
      src += sstep;
      dst += dstep;
      --left;
    }
  }
  __device__ __forceinline__ void rest() {
    if (!armed) return;
    while (left > 0) one();
    asm volatile("cp.async.commit_group;" ::: "memory");
    armed = false;
  }
};

// FETCH: how the next tile travels - 0: 16-byte cp.async pieces issued by every thread, 1: tensor boxes
// (cp.async.bulk.tensor, one thread), 2: one bulk copy per row (cp.async.bulk, W * 8 bytes each, rows spread over the
// block's threads; mbarrier completion like the boxes, no tensor map - the z sweep's rows lie a whole plane apart)
template <int M, int W, bool FINAL, int FETCH, bool BIG, bool UT, bool AHEAD>
__global__ void __launch_bounds__(BIG ? 512 : 256, BIG ? 1 : 2)
strided_sweep_tma(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ UTab ut,
                  const uint8_t *__restrict__ ucode, double *__restrict__ data, const double *__restrict__ Tin,
                  double *__restrict__ Tout, const uint32_t *__restrict__ line_id, const double *__restrict__ tab,
                  const double *__restrict__ GE, int L, int pitch, int P, int band, int64_t stride, int tiles_per_group,
                  int lines_per_group, int64_t group_stride, int n_tiles, int BR, int n_boxes, int do_prefetch, int pace) {
  extern __shared__ __align__(128) unsigned char smraw[];
  double *tile = reinterpret_cast<double *>(smraw);              // [n_boxes*BR][W]
  double *Y = tile + (size_t)n_boxes * BR * W;                     // [2P][W]
  double *Es = Y + 2 * P * W;                                      // [P][W]
  uint64_t *bar = reinterpret_cast<uint64_t *>(Es + P * W);
  double *s_tab = reinterpret_cast<double *>(bar + 2);             // [HS2_T_PLANES][pitch] of one unique line
  double *s_ge = s_tab + HS2_T_PLANES * pitch;                     // [P][2P]
  constexpr bool USE_TMA = FETCH != 0;                             // completion through the mbarrier
  const int w = threadIdx.x;
  const int p = threadIdx.y;
  const bool leader = (w == 0 && p == 0);
  const int r0 = p * M;
  const int rows = min(M, L - r0);
  const bool full = rows == M;
  const uint32_t tile_bytes = (uint32_t)n_boxes * BR * W * sizeof(double);

  if (USE_TMA) {
    if (leader) {
      mbar_init(bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
  }
  // fetch tile t into shared memory: one thread programs the TMA engine, or
  // (cp.async variant) every thread copies its share in 16-byte pieces
  auto issue = [&](int t) {
    const int group = t / tiles_per_group;
    const int c0 = (t % tiles_per_group) * W;
    if (FETCH == 1) {
      if (!leader) return;
      mbar_expect_tx(bar, tile_bytes);
      for (int b = 0; b < n_boxes; ++b) {
        if (group_stride)   // y-sweep: (x, row, plane)
          tma_load_3d(tile + (size_t)b * BR * W, &tmap, bar, c0, b * BR, group);
        else                // z-sweep: (flattened line, row, 0)
          tma_load_3d(tile + (size_t)b * BR * W, &tmap, bar, c0, b * BR, 0);
      }
    } else if (FETCH == 2) {
      const int nthr = W * P;
      const int tid = threadIdx.y * W + w;
      const uint32_t row_bytes = (uint32_t)min(W, lines_per_group - c0) * 8u;
      if (leader) mbar_expect_tx(bar, row_bytes * (uint32_t)L);
      const double *src = data + (int64_t)group * group_stride + c0 + (int64_t)tid * stride;
      uint32_t dst = smem_u32(tile) + (uint32_t)tid * (W * 8);
      const int64_t sstep = (int64_t)nthr * stride;
      for (int r = tid; r < L; r += nthr) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(src), "r"(row_bytes), "r"(smem_u32(bar))
                     : "memory");
        src += sstep;
        dst += (uint32_t)nthr * (W * 8);
      }
    } else {
      // thread tid copies the 16-byte piece (tid % (W/2)) of rows tid / (W/2), + 2 nthr / W, ...
      CpAsyncPacer pc;
      pc.setup(data + (int64_t)group * group_stride + c0, tile, threadIdx.y * W + w, W, W * P, L, stride, lines_per_group - c0);
      pc.rest();
    }
  };
  int t = blockIdx.x;
  if (t < n_tiles) issue(t);
  HS2_MARK_DECL;
  // factor tables of the first line's class into shared memory: persistent
  // blocks re-use them for every tile whose lines are of that class (all of
  // them on the BASELINE grids); other classes read the global tables
  uint32_t lid_c = 0xffffffffu;
  if (t < n_tiles) {
    lid_c = line_id[(int64_t)(t / tiles_per_group) * lines_per_group + (t % tiles_per_group) * W];
    const int tid = threadIdx.y * W + w, nthr = W * P;
    const double *gt = tab + (int64_t)lid_c * HS2_T_PLANES * pitch;
    for (int e = tid; e < HS2_T_PLANES * pitch; e += nthr) s_tab[e] = gt[e];
    const double *gg = GE + (int64_t)lid_c * P * 2 * P;
    for (int e = tid; e < 2 * P * P; e += nthr) s_ge[e] = gg[e];
  }
  __syncthreads();
  uint32_t parity = 0;
  // AHEAD: the unique-line id of this thread's line (and, in the z sweep, the common-table code that depends on it)
  // are fetched one tile ahead, so that the two dependent global loads are off the critical path of the tile.
  // Measured on B200: z sweep 0.634 -> 0.571 ms at 512^3 together with the lean T_in addressing; y sweep with several
  // line classes (steelonwater 512^3) 0.444 -> 0.411 ms; but the y sweep of one-class grids LOSES (the two registers
  // push the 128-register kernel into spills: 256-row lines 0.394 -> 0.438 ms, 1024-row lines 0.367 -> 0.434 ms), so
  // the launcher sets AHEAD for the z sweep and for y sweeps over more than one line class only.
  auto line_of = [&](int tt) {
    const int cc = (tt % tiles_per_group) * W + w;
    return (int64_t)(tt / tiles_per_group) * lines_per_group + (cc < lines_per_group ? cc : 0);
  };
  const bool has_ucode = UT && ucode != nullptr;
  uint32_t lid_n = (AHEAD && t < n_tiles) ? line_id[line_of(t)] : 0u;
  uint8_t uc_n = (AHEAD && has_ucode && t < n_tiles) ? ucode[(int64_t)lid_n * P + p] : (uint8_t)0;
  for (; t < n_tiles; t += gridDim.x) {
    const int group = t / tiles_per_group;
    const int col = (t % tiles_per_group) * W + w;
    const bool live = col < lines_per_group;
    const int64_t off = (int64_t)group * group_stride + (live ? col : 0) + (int64_t)r0 * stride;
    const uint32_t lid = AHEAD ? lid_n : line_id[line_of(t)];
    // every chunk of this warp carries the axis' common table: factors come from the constant bank
    const bool uni = has_ucode && __all_sync(0xffffffffu, (AHEAD ? uc_n : ucode[(int64_t)lid * P + p]) != 0);
    const bool more = t + (int)gridDim.x < n_tiles;
    if (AHEAD && more) lid_n = line_id[line_of(t + gridDim.x)];
    const double *tb = lid == lid_c ? s_tab + r0 : tab + ((int64_t)lid * HS2_T_PLANES) * pitch + r0;
    const double *ge = lid == lid_c ? s_ge + p * (2 * P) : GE + ((int64_t)lid * P + p) * (2 * P);
    const int64_t sbytes = stride * (int64_t)sizeof(double);
    if (FINAL && live && do_prefetch && (w & 3) == 0) {
      // warm L2 with the first do_prefetch T_in rows of the chunk (one lane per 32-byte sector); addresses by pointer
      // increments: base + q * stride with a 64-bit run-time stride cost ~20 instructions per row
      const char *ti = reinterpret_cast<const char *>(Tin + off);
      const int nq = min(rows, do_prefetch);
      if (nq == M) {
#pragma unroll
        for (int q = 0; q < M; ++q) {
          prefetch_l2(ti);
          ti += sbytes;
        }
      } else {
        for (int q = 0; q < nq; ++q) {
          prefetch_l2(ti);
          ti += sbytes;
        }
      }
    }
    HS2_MARK(8);
    if (USE_TMA) {
      mbar_wait(bar, parity);
      parity ^= 1;
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    }
    HS2_MARK(0);
    double v[M];
    {
      const double *mine = tile + (size_t)r0 * W + w;
#pragma unroll
      for (int q = 0; q < M; ++q) v[q] = (q < rows) ? mine[q * W] : 0.0;
    }
    HS2_MARK(1);
    __syncthreads();                       // tile buffer is free again
    // next tile: the copy engines get it at once; the cp.async pieces either in one go or (pace) one piece per
    // two rows of the forward elimination's dependent chain - free issue slots, and no burst of 4096 pieces per
    // block queueing ahead of the other resident block's T_in loads
    CpAsyncPacer pacer;
    if (more) {
      if (FETCH == 0 && pace) {
        const int tn = t + gridDim.x;
        const int c0n = (tn % tiles_per_group) * W;
        pacer.setup(data + (int64_t)(tn / tiles_per_group) * group_stride + c0n, tile, threadIdx.y * W + w, W, W * P, L, stride,
                    lines_per_group - c0n);
      } else {
        issue(t + gridDim.x);
      }
    }
    HS2_MARK(2);
    // (16-byte table loads: the 8-byte variant of chunk_core.cuh measured 6 % slower here)
    double yf, last;
    if (UT && uni) {
      yf = chunk_forward_const<M>(v, ut, pacer);
      last = v[M - 1];
    } else if (full) {
      yf = chunk_forward_full<M>(v, tb, pitch);
      last = v[M - 1];
    } else {
      yf = chunk_forward_short<M>(v, tb, pitch, rows, &last);
    }
    pacer.rest();
    Y[(2 * p) * W + w] = yf;
    Y[(2 * p + 1) * W + w] = last;
    HS2_MARK(3);
    __syncthreads();
    const double E = chunk_interface(ge, Y, P, W, w, p, band);
    Es[p * W + w] = E;
    if (AHEAD && more && has_ucode) uc_n = ucode[(int64_t)lid_n * P + p];       // lid_n arrived long ago
    HS2_MARK(4);
    __syncthreads();
    const double alpha = p > 0 ? Es[(p - 1) * W + w] : 0.0;
    if (UT && uni)
      chunk_backward_const<M>(v, ut, alpha, E);
    else if (full)
      chunk_backward_full<M>(v, tb, pitch, alpha, E);
    else
      chunk_backward_short<M>(v, tb, pitch, rows, alpha, E);
    HS2_MARK(5);
    if (live) {
      if (FINAL) {
        // T_in in batches of 8 rows, software-pipelined: the loads of batch g+1
        // are in flight while batch g is added and stored
        // (requesting the first batch before the back substitution measured slower: 0.68 vs 0.63 ms; a rolling window
        // of 12 / 16 loads that grows as rows are stored: 0.575 / 0.579 vs 0.572 ms at 512^3, 0.073 vs 0.078 at 256^3)
        double tin[2][8];
        if (full) {
          const char *ti = reinterpret_cast<const char *>(Tin + off);
          char *to = reinterpret_cast<char *>(Tout + off);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            tin[0][q] = *reinterpret_cast<const double *>(ti);
            ti += sbytes;
          }
#pragma unroll
          for (int g = 0; g < M; g += 8) {
            if (g + 8 < M) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                tin[((g >> 3) + 1) & 1][q] = *reinterpret_cast<const double *>(ti);
                ti += sbytes;
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              *reinterpret_cast<double *>(to) = tin[(g >> 3) & 1][q] + v[g + q];
              to += sbytes;
            }
          }
        } else {
          const double *ti = Tin + off;
          double *to = Tout + off;
#pragma unroll
          for (int q = 0; q < 8; ++q) tin[0][q] = (q < rows) ? ti[(int64_t)q * stride] : 0.0;
#pragma unroll
          for (int g = 0; g < M; g += 8) {
            if (g + 8 < M) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                tin[((g >> 3) + 1) & 1][q] = (g + 8 + q < rows) ? ti[(int64_t)(g + 8 + q) * stride] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (g + q < rows) to[(int64_t)(g + q) * stride] = tin[(g >> 3) & 1][q] + v[g + q];
          }
        }
      } else {
        double *dst = data + off;
#pragma unroll
        for (int q = 0; q < M; ++q) {
          if (q < rows) *dst = v[q];
          dst += stride;
        }
      }
    }
    HS2_MARK(6);
    // Y/Es are rewritten only after the next tile's first __syncthreads
  }
}

template <int M, bool FINAL>
int launch_tma(hs2_plan *pl, const hs2_axis_tables &ax, double *data, const double *Tin, double *Tout, int L,
               int64_t stride, int n_groups, int lines_per_group, int64_t group_stride, cudaStream_t st, bool *done) {
  *done = false;
  const int axis = FINAL ? 2 : 1;
  const int P = ax.n_chunks;
  constexpr int W = 16;
  // up to 16 chunks: 256-thread blocks, two per SM; up to 32 chunks (lines of
  // 513..1024 rows): one 512-thread block per SM with the whole 128 KB tile
  if (M < 32 && P * W > 256) return HS2_OK;
  if (P * W > 512) return HS2_OK;
  const bool big = P * W > 256;
  static const bool disabled = getenv("HS2_NO_TMA") != nullptr && getenv("HS2_NO_TMA")[0] == '1';
  if (disabled) return HS2_OK;
  // z-lines put consecutive rows 8*ny*nx bytes apart (a different 2 MB page per
  // row); measured on B200 the tensor copy engine then delivers < half the
  // bandwidth of plain loads (profiles/NOTES_r01.md), so the z-sweep keeps the
  // register-load kernel unless HS2_TMA_Z=1.
  // HS2_Z_PREFETCH: 0 = register-load kernel, 1 = cp.async persistent kernel (default), 2 = TMA boxes, 3 = bulk rows
  static const int zmode = getenv("HS2_Z_PREFETCH") ? atoi(getenv("HS2_Z_PREFETCH")) : 1;
  if (group_stride == 0 && zmode == 0) return HS2_OK;
  const bool use_tma = group_stride != 0 || zmode == 2;
  const bool use_rows = group_stride == 0 && zmode == 3 && !(lines_per_group & 1) && !(stride & 1) &&
                        !(reinterpret_cast<uintptr_t>(data) & 15);
  const uint8_t *ucode = (FINAL && pl->has_utab[axis] && !(pl->d.flags & HS2_FLAG_NO_UTAB)) ? ax.d_ucode : nullptr;
  // L2 warm-up of the tile's T_in rows at tile start (HS2_PREFETCH = rows per chunk to warm, 1 = all): off since the T_in / store phase addresses its rows by pointer
  // increments (measured at 512^3: z sweep 0.630 ms with it, 0.571 ms without; 256^3: 0.080 / 0.078)
  static const int pf_env = getenv("HS2_PREFETCH") ? atoi(getenv("HS2_PREFETCH")) : 0;
  const int pf = pf_env == 1 ? M : pf_env;
  static const int pace = getenv("HS2_Z_PACE") ? atoi(getenv("HS2_Z_PACE")) : 1;
  const int BR = L < 256 ? L : 256;
  const int n_boxes = (L + BR - 1) / BR;
  const size_t smem = ((size_t)n_boxes * BR * W + 3 * (size_t)P * W + (size_t)HS2_T_PLANES * ax.pitch + 2 * (size_t)P * P) * sizeof(double) + 16;
  if (smem > (big ? 224u : 112u) * 1024) return HS2_OK;
  const int tiles_per_group = (lines_per_group + W - 1) / W;
  const int64_t n_tiles = (int64_t)n_groups * tiles_per_group;
  if (n_tiles >= ((int64_t)1 << 31)) return HS2_OK;
  // resident blocks per SM: 16 warps in all (128 registers per thread), i.e. two 256-thread blocks, four of 128
  // threads (lines of 256 rows in chunks of 32) ... as far as the tiles fit shared memory
  static const int bps_env = getenv("HS2_STRIDED_BPS") ? atoi(getenv("HS2_STRIDED_BPS")) : 0;
  int bps = big ? 1 : 512 / (P * W);
  if (bps > 4) bps = 4;
  if (bps_env > 0 && !big) bps = bps_env;
  while (bps > 1 && (size_t)bps * (smem + 1024) > 226u * 1024) --bps;
  int grid = pl->sm_count * bps;
  if (grid > n_tiles) grid = (int)n_tiles;
  dim3 block(W, P);
  static const int carveout_env = getenv("HS2_CARVEOUT") ? atoi(getenv("HS2_CARVEOUT")) : -1;
  const int carveout = carveout_env >= 0 ? carveout_env : (int)((bps * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  // (the last template flag, AHEAD: see the kernel - always in the z sweep, in the y sweep for several line classes)
  auto kern = big ? strided_sweep_tma<M, W, FINAL, 0, (M >= 32), FINAL, FINAL> : strided_sweep_tma<M, W, FINAL, 0, false, FINAL, FINAL>;
  bool rows = false;
  if constexpr (FINAL) {
    if (use_rows) {
      kern = big ? strided_sweep_tma<M, W, FINAL, 2, (M >= 32), FINAL, FINAL> : strided_sweep_tma<M, W, FINAL, 2, false, FINAL, FINAL>;
      rows = true;
    }
  }
  if (rows) {
    pl->last_kernel[axis] = big ? HS2_K_TILE_ROWS_BIG : HS2_K_TILE_ROWS;
  } else if (use_tma) {
    bool ok;
    if (group_stride)
      ok = hs2_encode_tmap_f64_3d(&tmap, data, (uint64_t)lines_per_group, (uint64_t)L, (uint64_t)n_groups, W, BR, 1);
    else
      ok = hs2_encode_tmap_f64_3d(&tmap, data, (uint64_t)lines_per_group, (uint64_t)L, 1, W, BR, 1);
    if (!ok) return HS2_OK;
    kern = big ? strided_sweep_tma<M, W, FINAL, 1, (M >= 32), FINAL, FINAL> : strided_sweep_tma<M, W, FINAL, 1, false, FINAL, FINAL>;
    if constexpr (!FINAL) {
      static const int ahead_env = getenv("HS2_Y_AHEAD") ? atoi(getenv("HS2_Y_AHEAD")) : -1;
      const bool ahead = ahead_env >= 0 ? ahead_env != 0 : (ax.n_unique > 1 && P == 16);
      if (ahead && !big) kern = strided_sweep_tma<M, W, FINAL, 1, false, FINAL, true>;
    }
    pl->last_kernel[axis] = big ? HS2_K_TILE_TMA_BIG : HS2_K_TILE_TMA;
  } else {
    // 16-byte cp.async pieces: every row start must be 16-byte aligned
    if ((lines_per_group & 1) || (stride & 1) || (group_stride & 1) || (reinterpret_cast<uintptr_t>(data) & 15)) return HS2_OK;
    pl->last_kernel[axis] = big ? HS2_K_TILE_CPASYNC_BIG : HS2_K_TILE_CPASYNC;
  }
  HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // two blocks per SM: leave the rest of the 256 KB array to L1 (factor tables live there)
  HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout > 100 ? 100 : carveout));
  kern<<<grid, block, smem, st>>>(tmap, pl->utab[axis], ucode, data, Tin, Tout, ax.d_line_id, ax.d_tab, ax.d_GE, L, ax.pitch, P, ax.band,
                                  stride, tiles_per_group, lines_per_group, group_stride, (int)n_tiles, BR, n_boxes, pf, pace);
  HS2_CUDA_CHECK(cudaGetLastError());
  *done = true;
  return HS2_OK;
}

template <int M, bool FINAL>
int launch(const hs2_axis_tables &ax, double *data, const double *Tin, double *Tout, int L, int64_t stride,
           int n_groups, int lines_per_group, int64_t group_stride, cudaStream_t st) {
  const int P = ax.n_chunks;
  const int maxthreads = M >= 32 ? 256 : 512;
  const int W = P * 16 <= maxthreads ? 16 : 8;
  HS2_REQUIRE(P * W <= maxthreads, "strided sweep: %d chunks of %d rows do not fit a block", P, M);
  const int tiles_per_group = (lines_per_group + W - 1) / W;
  const int64_t blocks = (int64_t)n_groups * tiles_per_group;
  HS2_REQUIRE(blocks < ((int64_t)1 << 31), "strided sweep: too many tiles");
  dim3 block(W, P);
  const size_t smem = (size_t)3 * P * W * sizeof(double);
  if (W == 16)
    strided_sweep<M, 16, FINAL><<<(unsigned)blocks, block, smem, st>>>(data, Tin, Tout, ax.d_line_id, ax.d_tab, ax.d_GE, L,
                                                                      ax.pitch, P, ax.band, stride, tiles_per_group,
                                                                      lines_per_group, group_stride);
  else
    strided_sweep<M, 8, FINAL><<<(unsigned)blocks, block, smem, st>>>(data, Tin, Tout, ax.d_line_id, ax.d_tab, ax.d_GE, L,
                                                                     ax.pitch, P, ax.band, stride, tiles_per_group,
                                                                     lines_per_group, group_stride);
  HS2_CUDA_CHECK(cudaGetLastError());
  return HS2_OK;
}

template <bool FINAL>
int dispatch(hs2_plan *pl, const hs2_axis_tables &ax, double *data, const double *Tin, double *Tout, int L,
             int64_t stride, int n_groups, int lines_per_group, int64_t group_stride, cudaStream_t st) {
  bool done = false;
  int rc = HS2_OK;
  switch (ax.chunk) {
    case 8: rc = launch_tma<8, FINAL>(pl, ax, data, Tin, Tout, L, stride, n_groups, lines_per_group, group_stride, st, &done); break;
    case 16: rc = launch_tma<16, FINAL>(pl, ax, data, Tin, Tout, L, stride, n_groups, lines_per_group, group_stride, st, &done); break;
    case 32: rc = launch_tma<32, FINAL>(pl, ax, data, Tin, Tout, L, stride, n_groups, lines_per_group, group_stride, st, &done); break;
  }
  if (rc || done) return rc;
  pl->last_kernel[FINAL ? 2 : 1] = HS2_K_TILE;
  switch (ax.chunk) {
    case 8: return launch<8, FINAL>(ax, data, Tin, Tout, L, stride, n_groups, lines_per_group, group_stride, st);
    case 16: return launch<16, FINAL>(ax, data, Tin, Tout, L, stride, n_groups, lines_per_group, group_stride, st);
    case 32: return launch<32, FINAL>(ax, data, Tin, Tout, L, stride, n_groups, lines_per_group, group_stride, st);
  }
  hs2_set_error("strided sweep: unsupported chunk size %d", ax.chunk);
  return HS2_E_INVALID;
}

}  // namespace

// ---------------------------------------------------------------------------
// z-sweep of a z-slab decomposition (one slab per GPU).  The global z-line is
// cut into chunks of M rows; a slab owns P_loc consecutive chunks
// [chunk0, chunk0 + P_loc).  Phase 1 (z_forward) eliminates every local chunk
// and publishes its (y_first, y_last); the host layer all-gathers those 2
// doubles per chunk per line over NCCL; phase 2 (z_backward) applies the rows
// p-1 and p of the GLOBAL inverse interface operator to get alpha and E and
// back-substitutes.  No transpose of the field is ever needed.
struct ZPeers {  // where this slab's (y_first, y_last) rows go in the peers' interface buffers
  int n;
  double *y[HS2_MAX_Z_PEERS];
};

// Both kernels: block = (W lines, P_loc chunks, G line groups) = 256 threads,
// persistent over groups of G tiles.  The factor-table rows of this slab and the
// P_loc+1 rows of the global inverse interface operator it needs are copied to
// shared memory once per block (class of the block's first line; lines of
// another class read global memory) and read with 8-byte broadcast loads.
template <int M, int W>
__global__ void __launch_bounds__(256, 2)
z_forward(const double *__restrict__ data, const uint32_t *__restrict__ line_id, const double *__restrict__ tab,
          double *__restrict__ Yloc, int pitch, int row0, int nz_loc, int64_t stride, int line0, int n_lines, int n_tiles,
          ZPeers peers, int64_t ldy, int ycol0) {
  extern __shared__ __align__(16) double zsm[];
  double *s_tab = zsm;                                   // [HS2_T_PLANES][nz_loc]
  const int w = threadIdx.x, p = threadIdx.y, g = threadIdx.z, G = blockDim.z;
  const int tid = (g * blockDim.y + p) * W + w, nthr = W * blockDim.y * G;
  int tile = blockIdx.x * G;
  if (tile >= n_tiles) return;
  const uint32_t lid_c = line_id[line0 + tile * W];
  for (int e = tid; e < HS2_T_PLANES * nz_loc; e += nthr)
    s_tab[e] = tab[((int64_t)lid_c * HS2_T_PLANES + e / nz_loc) * pitch + row0 + e % nz_loc];
  __syncthreads();
  TabShared ts;
  ts.a = smem_u32(s_tab + p * M);
  ts.pitch_b = (uint32_t)nz_loc * 8u;
  for (; tile < n_tiles; tile += gridDim.x * G) {
    const int rel = (tile + g) * W + w;          // line within the processed range
    if (tile + g >= n_tiles || rel >= n_lines) continue;
    const int col = line0 + rel;
    const uint32_t lid = line_id[col];
    const char *ptr = reinterpret_cast<const char *>(data + col + (int64_t)p * M * stride);
    const int64_t sbytes = stride * (int64_t)sizeof(double);     // rows by pointer increments (no 64-bit multiply per row)
    double v[M];
#pragma unroll
    for (int t = 0; t < M; ++t) {
      v[t] = *reinterpret_cast<const double *>(ptr);
      ptr += sbytes;
    }
    // the eliminated chunk is NOT written back: z_backward repeats the (cheap)
    // forward elimination from the same input instead of re-reading 8 B/cell
    double yf, last;
    if (lid == lid_c) {
      yf = chunk_fwd<M, true>(v, ts, M, &last);
    } else {
      TabGlobal tg;
      tg.b = tab + ((int64_t)lid * HS2_T_PLANES) * pitch + row0 + p * M;
      tg.pitch = pitch;
      yf = chunk_fwd<M, true>(v, tg, M, &last);
    }
    // interface rows: [2 P][ldy], this launch's lines at columns ycol0 + rel
    const int64_t y0 = (int64_t)(2 * p) * ldy + ycol0 + rel, y1 = y0 + ldy;
    Yloc[y0] = yf;
    Yloc[y1] = last;
    // the same two values straight into the memory of the slabs that need them (NVLink stores)
#pragma unroll
    for (int q = 0; q < HS2_MAX_Z_PEERS; ++q) {
      if (q < peers.n) {
        peers.y[q][y0] = yf;
        peers.y[q][y1] = last;
      }
    }
  }
}

template <int M, int W>
__global__ void __launch_bounds__(256, 2)
z_backward(const double *__restrict__ data, const double *__restrict__ Tin, double *__restrict__ Tout,
           const uint32_t *__restrict__ line_id, const double *__restrict__ tab, const double *__restrict__ GE,
           const double *__restrict__ Yall, int pitch, int row0, int nz_loc, int P_glob, int chunk0, int band,
           int64_t stride, int line0, int n_lines, int n_tiles, int64_t ldy, int ycol0) {
  extern __shared__ __align__(16) double zsm[];
  double *s_tab = zsm;                                   // [HS2_T_PLANES][nz_loc]
  double *s_ge = s_tab + HS2_T_PLANES * nz_loc;          // [P_loc + 1][2 P_glob]: rows chunk0-1 .. chunk0+P_loc-1
  const int w = threadIdx.x, p = threadIdx.y, g = threadIdx.z, G = blockDim.z;
  const int P_loc = blockDim.y;
  const int tid = (g * P_loc + p) * W + w, nthr = W * P_loc * G;
  int tile = blockIdx.x * G;
  if (tile >= n_tiles) return;
  const uint32_t lid_c = line_id[line0 + tile * W];
  for (int e = tid; e < HS2_T_PLANES * nz_loc; e += nthr)
    s_tab[e] = tab[((int64_t)lid_c * HS2_T_PLANES + e / nz_loc) * pitch + row0 + e % nz_loc];
  for (int e = tid; e < (P_loc + 1) * 2 * P_glob; e += nthr) {
    const int row = chunk0 - 1 + e / (2 * P_glob);
    s_ge[e] = row >= 0 ? GE[((int64_t)lid_c * P_glob + row) * (2 * P_glob) + e % (2 * P_glob)] : 0.0;
  }
  __syncthreads();
  const int pg = chunk0 + p;
  TabShared ts;
  ts.a = smem_u32(s_tab + p * M);
  ts.pitch_b = (uint32_t)nz_loc * 8u;
  const uint32_t ge_s = smem_u32(s_ge + (p + 1) * 2 * P_glob);   // row pg; row pg-1 sits just before it
  for (; tile < n_tiles; tile += gridDim.x * G) {
    const int rel = (tile + g) * W + w;
    if (tile + g >= n_tiles || rel >= n_lines) continue;
    const int col = line0 + rel;
    const uint32_t lid = line_id[col];
    const bool tab_s = lid == lid_c;
    const int64_t off = col + (int64_t)p * M * stride;
    const int64_t sbytes = stride * (int64_t)sizeof(double);     // rows by pointer increments (no 64-bit multiply per row)
    double v[M];
    {
      const char *ptr = reinterpret_cast<const char *>(data + off);
#pragma unroll
      for (int t = 0; t < M; ++t) {
        v[t] = *reinterpret_cast<const double *>(ptr);
        ptr += sbytes;
      }
    }
    // rows pg (-> E) and pg-1 (-> alpha) of the inverse interface operator
    double E = 0.0, alpha = 0.0;
    const double *Yc = Yall + ycol0 + rel;               // this line's column of the interface rows [2 P_glob][ldy]
    {
      const double *ge = GE + ((int64_t)lid * P_glob + pg) * (2 * P_glob);
      const int q0 = max(0, pg - band), q1 = min(P_glob - 1, pg + band);
      for (int q = q0; q <= q1; ++q) {
        const double g0 = tab_s ? TabShared::ld(ge_s, 2 * q) : __ldg(ge + 2 * q);
        const double g1 = tab_s ? TabShared::ld(ge_s, 2 * q + 1) : __ldg(ge + 2 * q + 1);
        E = fma(g0, Yc[(int64_t)(2 * q) * ldy], E);
        E = fma(g1, Yc[(int64_t)(2 * q + 1) * ldy], E);
      }
      if (pg > 0) {
        const double *gm = ge - 2 * P_glob;
        const uint32_t gm_s = ge_s - 16u * (uint32_t)P_glob;
        const int a0 = max(0, pg - 1 - band), a1 = min(P_glob - 1, pg - 1 + band);
        for (int q = a0; q <= a1; ++q) {
          const double g0 = tab_s ? TabShared::ld(gm_s, 2 * q) : __ldg(gm + 2 * q);
          const double g1 = tab_s ? TabShared::ld(gm_s, 2 * q + 1) : __ldg(gm + 2 * q + 1);
          alpha = fma(g0, Yc[(int64_t)(2 * q) * ldy], alpha);
          alpha = fma(g1, Yc[(int64_t)(2 * q + 1) * ldy], alpha);
        }
      }
    }
    double last;
    if (tab_s) {
      chunk_fwd<M, true>(v, ts, M, &last);       // same arithmetic as z_forward
      chunk_bwd<M, true>(v, ts, M, alpha, E);
    } else {
      TabGlobal tg;
      tg.b = tab + ((int64_t)lid * HS2_T_PLANES) * pitch + row0 + p * M;
      tg.pitch = pitch;
      chunk_fwd<M, true>(v, tg, M, &last);
      chunk_bwd<M, true>(v, tg, M, alpha, E);
    }
    // T_in in batches of 8 rows, software-pipelined: the loads of batch g+1 are
    // in flight while batch g is added and stored
    double tin[2][8];
    const char *ti = reinterpret_cast<const char *>(Tin + off);
    char *to = reinterpret_cast<char *>(Tout + off);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      tin[0][q] = *reinterpret_cast<const double *>(ti);
      ti += sbytes;
    }
#pragma unroll
    for (int gb = 0; gb < M; gb += 8) {
      if (gb + 8 < M) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          tin[((gb >> 3) + 1) & 1][q] = *reinterpret_cast<const double *>(ti);
          ti += sbytes;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        *reinterpret_cast<double *>(to) = tin[(gb >> 3) & 1][q] + v[gb + q];
        to += sbytes;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Fused slab z sweep (peer-memory transport): forward elimination, exchange of the interface rows and back
// substitution in ONE persistent kernel, so that the eliminated chunk never leaves the chip (z_forward +
// z_backward read the increment twice: 32 instead of 24 B per cell) and no wait is exposed between two launches.
//
// The interface rows are their own flags: every slot of a mailbox holds HS2_Y_EMPTY (a NaN payload no arithmetic
// produces) until the producing slab stores the value - one naturally aligned 8-byte store over NVLink, never torn -
// and the consumer puts HS2_Y_EMPTY back once the block has read it (the rows are double-buffered by step parity and
// a slab is never two steps ahead of a neighbour, dist.py).  No fence, no flag, no thread that publishes for the block.
// (A first version with one flag per tile - barrier, fence.sys and release store by one thread, acquire spin by one
// thread - ran at 1.4 ms per 512^3 slab against 0.84 ms for z_forward + z_backward: profiles/NOTES_r02.md.)
// Per tile (the 256 threads' lines: W lines x G groups, all local chunks):
//   forward elimination from registers; (y_first, y_last) stored into the own rows and straight into the peers'
//   mailboxes; the eliminated chunk is parked in shared memory while the NEXT tile is eliminated and sent;
//   then the tile comes back into registers, every thread reads the 2 (2 band + 2) interface values of its chunk
//   from L2 (ld.cg), spinning (bounded) on the ones that are still empty; barrier; the peers' slots are emptied;
//   back substitution, T_in + increment.
// A block handles its tiles in increasing order and never waits for a tile before it has sent that tile and the
// next one, so two slabs cannot wait for each other, whatever the number of resident blocks on either GPU.
#define HS2_Y_EMPTY 0x7FF8DEADBEEF5A5Aull

__global__ void fill_u64_kernel(unsigned long long *p, int64_t n, unsigned long long v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

template <int M, int W>
__global__ void __launch_bounds__(256, 2)
z_fused(const double *__restrict__ data, const double *__restrict__ Tin, double *__restrict__ Tout,
        const uint32_t *__restrict__ line_id, const double *__restrict__ tab, const double *__restrict__ GE,
        double *Yall, int pitch, int row0, int nz_loc, int P_glob, int chunk0, int band, int64_t stride, int n_lines,
        int n_tiles, ZPeers peers, long long max_cycles, int *status) {
  extern __shared__ __align__(16) double zsm[];
  double *s_tab = zsm;                                   // [HS2_T_PLANES][nz_loc]
  double *s_ge = s_tab + HS2_T_PLANES * nz_loc;          // [P_loc + 1][2 P_glob]: rows chunk0-1 .. chunk0+P_loc-1
  const int w = threadIdx.x, p = threadIdx.y, g = threadIdx.z, G = blockDim.z;
  const int P_loc = blockDim.y;
  double *s_park = s_ge + (P_loc + 1) * 2 * P_glob;      // [M][256]: the eliminated chunk of the tile that waits
  const int tid = (g * P_loc + p) * W + w, nthr = W * P_loc * G;
  int tile = blockIdx.x * G;
  if (tile >= n_tiles) return;
  const uint32_t lid_c = line_id[tile * W];
  for (int e = tid; e < HS2_T_PLANES * nz_loc; e += nthr)
    s_tab[e] = tab[((int64_t)lid_c * HS2_T_PLANES + e / nz_loc) * pitch + row0 + e % nz_loc];
  for (int e = tid; e < (P_loc + 1) * 2 * P_glob; e += nthr) {
    const int row = chunk0 - 1 + e / (2 * P_glob);
    s_ge[e] = row >= 0 ? GE[((int64_t)lid_c * P_glob + row) * (2 * P_glob) + e % (2 * P_glob)] : 0.0;
  }
  __syncthreads();
  const int pg = chunk0 + p;
  // the threads of line group g (W lines x all local chunks: whole warps) only ever wait for each other
  // (named barriers: whole warps only, ids 1..15)
  const bool own_bar = G > 1 && G <= 15 && ((W * P_loc) & 31) == 0;
  auto group_sync = [&]() {
    if (!own_bar)
      __syncthreads();
    else
      asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(W * P_loc) : "memory");
  };
  TabShared ts;
  ts.a = smem_u32(s_tab + p * M);
  ts.pitch_b = (uint32_t)nz_loc * 8u;
  const uint32_t ge_s = smem_u32(s_ge + (p + 1) * 2 * P_glob);
  double *Yloc = Yall + (int64_t)(2 * chunk0) * n_lines;
  double *park = s_park + tid;
  bool dead = *reinterpret_cast<volatile int *>(status) != 0;     // an earlier wait timed out: do not wait again

  double v[M];
  // eliminate the thread's chunk of tile t0 (kept in v); its interface values go to the own rows and to the peers
  auto forward = [&](int t0) {
    const int rel = (t0 + g) * W + w;
    if (t0 + g < n_tiles && rel < n_lines) {
      const uint32_t lid = line_id[rel];
      const double *ptr = data + rel + (int64_t)p * M * stride;
#pragma unroll
      for (int t = 0; t < M; ++t) v[t] = ptr[(int64_t)t * stride];
      double yf, last;
      if (lid == lid_c) {
        yf = chunk_fwd<M, true>(v, ts, M, &last);
      } else {
        TabGlobal tg;
        tg.b = tab + ((int64_t)lid * HS2_T_PLANES) * pitch + row0 + p * M;
        tg.pitch = pitch;
        yf = chunk_fwd<M, true>(v, tg, M, &last);
      }
      const int64_t y0 = (int64_t)(2 * p) * n_lines + rel, y1 = y0 + n_lines;
      Yloc[y0] = yf;
      Yloc[y1] = last;
#pragma unroll
      for (int q = 0; q < HS2_MAX_Z_PEERS; ++q) {
        if (q < peers.n) {
          peers.y[q][y0] = yf;
          peers.y[q][y1] = last;
        }
      }
    }
  };
  // one interface value: own rows are ordered by the block barrier, a peer's slot is read until it is no longer empty
  const long long t_begin = clock64();
  auto y_get = [&](const double *q, bool own) {
    double x = __ldcg(q);
    if (!own && !dead) {
      unsigned ns = 256;                 // back off: tens of thousands of threads polling L2 every 64 ns starve the data traffic
      while (__double_as_longlong(x) == (long long)HS2_Y_EMPTY) {
        if (clock64() - t_begin > max_cycles) {
          atomicExch(status, 1);
          dead = true;
          break;
        }
        __nanosleep(ns);
        if (ns < 4096) ns <<= 1;
        x = __ldcg(q);
      }
    }
    return x;
  };

  // The phases of a tile (rows of d, forward, interface values, backward, rows of T_in, stores) follow each other
  // inside a block and only two blocks share an SM, so every row is asked for one tile ahead: one L2 prefetch per
  // 128-byte piece (the W lines of a row), issued by one thread per chunk
  auto warm = [&](const double *base, int t0) {
    const int rel0 = (t0 + g) * W;
    if (w == 0 && t0 + g < n_tiles && rel0 < n_lines) {
      const double *ptr = base + rel0 + (int64_t)p * M * stride;
#pragma unroll
      for (int t = 0; t < M; ++t) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + (int64_t)t * stride));
    }
  };
  warm(Tin, tile);
  warm(data, tile + gridDim.x * G);
  forward(tile);
#pragma unroll
  for (int t = 0; t < M; ++t) park[t * 256] = v[t];
  for (;;) {
    const int next = tile + gridDim.x * G;
    const bool has_next = next < n_tiles;
    if (has_next) {
      warm(Tin, next);
      warm(data, next + gridDim.x * G);
      forward(next);
#pragma unroll
      for (int t = 0; t < M; ++t) {      // swap: the parked tile comes back, the new one is parked (own slots only)
        const double x = park[t * 256];
        park[t * 256] = v[t];
        v[t] = x;
      }
    } else {
#pragma unroll
      for (int t = 0; t < M; ++t) v[t] = park[t * 256];
    }
    group_sync();                        // the own rows of the group's lines are there (all chunks of a line sit in one group)
    const int rel = (tile + g) * W + w;
    const bool live = tile + g < n_tiles && rel < n_lines;
    uint32_t lid = 0;
    double E = 0.0, alpha = 0.0;
    double *Yc = Yall + rel;
    if (live) {
      lid = line_id[rel];
      const bool tab_s = lid == lid_c;
      const double *ge = GE + ((int64_t)lid * P_glob + pg) * (2 * P_glob);
      const int q0 = max(0, pg - band), q1 = min(P_glob - 1, pg + band);
      for (int q = q0; q <= q1; ++q) {
        const bool own = q >= chunk0 && q < chunk0 + P_loc;
        const double g0 = tab_s ? TabShared::ld(ge_s, 2 * q) : __ldg(ge + 2 * q);
        const double g1 = tab_s ? TabShared::ld(ge_s, 2 * q + 1) : __ldg(ge + 2 * q + 1);
        E = fma(g0, y_get(Yc + (int64_t)(2 * q) * n_lines, own), E);
        E = fma(g1, y_get(Yc + (int64_t)(2 * q + 1) * n_lines, own), E);
      }
      if (pg > 0) {
        const double *gm = ge - 2 * P_glob;
        const uint32_t gm_s = ge_s - 16u * (uint32_t)P_glob;
        const int a0 = max(0, pg - 1 - band), a1 = min(P_glob - 1, pg - 1 + band);
        for (int q = a0; q <= a1; ++q) {
          const bool own = q >= chunk0 && q < chunk0 + P_loc;
          const double g0 = tab_s ? TabShared::ld(gm_s, 2 * q) : __ldg(gm + 2 * q);
          const double g1 = tab_s ? TabShared::ld(gm_s, 2 * q + 1) : __ldg(gm + 2 * q + 1);
          alpha = fma(g0, y_get(Yc + (int64_t)(2 * q) * n_lines, own), alpha);
          alpha = fma(g1, y_get(Yc + (int64_t)(2 * q + 1) * n_lines, own), alpha);
        }
      }
    }
    group_sync();                        // every chunk of the line has read the peers' slots
    if (live) {
      // empty the peers' slots this line read (chunks pg-1-band .. pg+band of the first local chunk, the new top
      // one of the others), for the step after next
      const int lo = p == 0 ? max(0, pg - 1 - band) : pg + band, hi = min(P_glob - 1, pg + band);
      for (int q = lo; q <= hi; ++q) {
        if (q >= chunk0 && q < chunk0 + P_loc) continue;
        reinterpret_cast<unsigned long long *>(Yc)[(int64_t)(2 * q) * n_lines] = HS2_Y_EMPTY;
        reinterpret_cast<unsigned long long *>(Yc)[(int64_t)(2 * q + 1) * n_lines] = HS2_Y_EMPTY;
      }
      const int64_t off = rel + (int64_t)p * M * stride;
      if (lid == lid_c) {
        chunk_bwd<M, true>(v, ts, M, alpha, E);
      } else {
        TabGlobal tg;
        tg.b = tab + ((int64_t)lid * HS2_T_PLANES) * pitch + row0 + p * M;
        tg.pitch = pitch;
        chunk_bwd<M, true>(v, tg, M, alpha, E);
      }
      double tin[2][8];
#pragma unroll
      for (int q = 0; q < 8; ++q) tin[0][q] = Tin[off + (int64_t)q * stride];
#pragma unroll
      for (int gb = 0; gb < M; gb += 8) {
        if (gb + 8 < M) {
#pragma unroll
          for (int q = 0; q < 8; ++q) tin[((gb >> 3) + 1) & 1][q] = Tin[off + (int64_t)(gb + 8 + q) * stride];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) Tout[off + (int64_t)(gb + q) * stride] = tin[(gb >> 3) & 1][q] + v[gb + q];
      }
    }
    if (!has_next) break;
    tile = next;
  }
}

// Warp-autonomous variant of the fused slab z sweep, for thin slabs: ALL chunks of a line sit in one warp
// (WL = 32 / P_loc lines x P_loc chunks, lane = p * WL + w; WL lines = WL * 8 contiguous bytes per row), so the
// interface values of the own chunks travel by shuffle, the peers' values are polled from the mailbox slots
// (self-validating, see z_fused), and there is no barrier at all: sixteen independent warps per SM, each in its own
// phase.  No parking either: while a warp waits for its peers' rows the other warps use the memory system.
template <int M, int WL>
__global__ void __launch_bounds__(256, 2)
z_fused_warp(const double *__restrict__ data, const double *__restrict__ Tin, double *__restrict__ Tout,
             const uint32_t *__restrict__ line_id, const double *__restrict__ tab, const double *__restrict__ GE,
             double *Yall, int pitch, int row0, int nz_loc, int P_glob, int chunk0, int band, int64_t stride, int n_lines,
             int n_tiles, ZPeers peers, long long max_cycles, int *status) {
  constexpr int P_loc = 32 / WL;
  extern __shared__ __align__(16) double zsm[];
  double *s_tab = zsm;                                   // [HS2_T_PLANES][nz_loc]
  double *s_ge = s_tab + HS2_T_PLANES * nz_loc;          // [P_loc + 1][2 P_glob]: rows chunk0-1 .. chunk0+P_loc-1
  const int tid = threadIdx.x, lane = tid & 31;
  const int w = lane % WL, p = lane / WL;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  int tile = (blockIdx.x * blockDim.x + tid) >> 5;
  const uint32_t lid_c = line_id[min((int64_t)((blockIdx.x * blockDim.x) >> 5) * WL, (int64_t)n_lines - 1)];
  for (int e = tid; e < HS2_T_PLANES * nz_loc; e += blockDim.x)
    s_tab[e] = tab[((int64_t)lid_c * HS2_T_PLANES + e / nz_loc) * pitch + row0 + e % nz_loc];
  for (int e = tid; e < (P_loc + 1) * 2 * P_glob; e += blockDim.x) {
    const int row = chunk0 - 1 + e / (2 * P_glob);
    s_ge[e] = row >= 0 ? GE[((int64_t)lid_c * P_glob + row) * (2 * P_glob) + e % (2 * P_glob)] : 0.0;
  }
  __syncthreads();
  const int pg = chunk0 + p;
  TabShared ts;
  ts.a = smem_u32(s_tab + p * M);
  ts.pitch_b = (uint32_t)nz_loc * 8u;
  const uint32_t ge_s = smem_u32(s_ge + (p + 1) * 2 * P_glob);      // row pg; row pg-1 sits just before it
  bool dead = *reinterpret_cast<volatile int *>(status) != 0;
  const long long t_begin = clock64();

  for (; tile < n_tiles; tile += warps) {
    const int rel = tile * WL + w;
    const bool live = rel < n_lines;
    const int relc = live ? rel : n_lines - 1;           // (lanes without a line go through the motions on the last one)
    const uint32_t lid = line_id[relc];
    const bool tab_s = lid == lid_c;
    const int64_t off = relc + (int64_t)p * M * stride;
    double v[M];
#pragma unroll
    for (int t = 0; t < M; ++t) v[t] = data[off + (int64_t)t * stride];
    double yf, last;
    TabGlobal tg;
    tg.b = tab + ((int64_t)lid * HS2_T_PLANES) * pitch + row0 + p * M;
    tg.pitch = pitch;
    if (tab_s)
      yf = chunk_fwd<M, true>(v, ts, M, &last);
    else
      yf = chunk_fwd<M, true>(v, tg, M, &last);
    if (live) {
      const int64_t y0 = (int64_t)(2 * p) * n_lines + rel, y1 = y0 + n_lines;
#pragma unroll
      for (int q = 0; q < HS2_MAX_Z_PEERS; ++q) {
        if (q < peers.n) {
          peers.y[q][y0] = yf;
          peers.y[q][y1] = last;
        }
      }
    }
    // interface values of the chunks pg-1-band .. pg+band: own chunks by shuffle, the peers' from their slots
    double E = 0.0, alpha = 0.0;
    double *Yc = Yall + relc;
    const double *ge = GE + ((int64_t)lid * P_glob + pg) * (2 * P_glob);
    for (int d = -band - 1; d <= band; ++d) {
      const int q = pg + d;
      const int src = lane + d * WL;
      const double ys = __shfl_sync(0xffffffffu, yf, src & 31), ls = __shfl_sync(0xffffffffu, last, src & 31);
      if (q < 0 || q >= P_glob) continue;
      double y_f = ys, y_l = ls;
      const bool own = q >= chunk0 && q < chunk0 + P_loc;
      if (!own) {
        const double *s0 = Yc + (int64_t)(2 * q) * n_lines, *s1 = s0 + n_lines;
        y_f = __ldcg(s0), y_l = __ldcg(s1);
        if (live && !dead) {
          unsigned ns = 256;             // back off: a warp's 32 lanes x 16 warps x 148 SMs polling every 64 ns flood L2
          while (__double_as_longlong(y_f) == (long long)HS2_Y_EMPTY || __double_as_longlong(y_l) == (long long)HS2_Y_EMPTY) {
            if (clock64() - t_begin > max_cycles) {
              atomicExch(status, 1);
              dead = true;
              break;
            }
            __nanosleep(ns);
            if (ns < 4096) ns <<= 1;
            y_f = __ldcg(s0), y_l = __ldcg(s1);
          }
        }
      }
      if (d >= -band) {      // row pg: chunks pg-band .. pg+band
        const double g0 = tab_s ? TabShared::ld(ge_s, 2 * q) : __ldg(ge + 2 * q);
        const double g1 = tab_s ? TabShared::ld(ge_s, 2 * q + 1) : __ldg(ge + 2 * q + 1);
        E = fma(g0, y_f, E);
        E = fma(g1, y_l, E);
      }
      if (pg > 0 && d < band) {      // row pg-1: chunks pg-1-band .. pg-1+band
        const double g0 = tab_s ? TabShared::ld(ge_s - 16u * (uint32_t)P_glob, 2 * q) : __ldg(ge - 2 * P_glob + 2 * q);
        const double g1 = tab_s ? TabShared::ld(ge_s - 16u * (uint32_t)P_glob, 2 * q + 1) : __ldg(ge - 2 * P_glob + 2 * q + 1);
        alpha = fma(g0, y_f, alpha);
        alpha = fma(g1, y_l, alpha);
      }
    }
    __syncwarp();                        // every chunk of the line has read the peers' slots
    if (live) {
      const int lo = p == 0 ? max(0, pg - 1 - band) : pg + band, hi = min(P_glob - 1, pg + band);
      for (int q = lo; q <= hi; ++q) {
        if (q >= chunk0 && q < chunk0 + P_loc) continue;
        reinterpret_cast<unsigned long long *>(Yc)[(int64_t)(2 * q) * n_lines] = HS2_Y_EMPTY;
        reinterpret_cast<unsigned long long *>(Yc)[(int64_t)(2 * q + 1) * n_lines] = HS2_Y_EMPTY;
      }
    }
    if (tab_s)
      chunk_bwd<M, true>(v, ts, M, alpha, E);
    else
      chunk_bwd<M, true>(v, tg, M, alpha, E);
    if (live) {
      double tin[2][8];
#pragma unroll
      for (int q = 0; q < 8; ++q) tin[0][q] = Tin[off + (int64_t)q * stride];
#pragma unroll
      for (int gb = 0; gb < M; gb += 8) {
        if (gb + 8 < M) {
#pragma unroll
          for (int q = 0; q < 8; ++q) tin[((gb >> 3) + 1) & 1][q] = Tin[off + (int64_t)(gb + 8 + q) * stride];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) Tout[off + (int64_t)(gb + q) * stride] = tin[(gb >> 3) & 1][q] + v[gb + q];
      }
    }
  }
}

template <int M, int WL>
int launch_zfused_warp(hs2_plan *pl, double *data, const double *Tin, double *Tout, double *Yall, const ZPeers &peers,
                       double timeout_s, int *status, int *tile_lines, cudaStream_t st) {
  const hs2_plan_desc &d = pl->d;
  const hs2_axis_tables &ax = d.axis[2];
  constexpr int P_loc = 32 / WL;
  if (tile_lines) {
    *tile_lines = WL;
    return HS2_OK;
  }
  const int n_lines = (int)(d.ny * d.nx);
  const int n_tiles = (n_lines + WL - 1) / WL;
  const int row0 = d.z_chunk0 * M;
  const int nz_loc = (int)d.nz;
  const size_t smem = ((size_t)HS2_T_PLANES * nz_loc + (size_t)(P_loc + 1) * 2 * d.z_chunks_global) * sizeof(double);
  HS2_REQUIRE(smem <= 100 * 1024, "fused z sweep: tables need %zu B of shared memory", smem);
  auto kern = z_fused_warp<M, WL>;
  if (smem > 48 * 1024) HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = pl->sm_count * 2;
  if (blocks > (n_tiles + 7) / 8) blocks = (n_tiles + 7) / 8;
  kern<<<blocks, 256, smem, st>>>(data, Tin, Tout, ax.d_line_id, ax.d_tab, ax.d_GE, Yall, ax.pitch, row0, nz_loc, d.z_chunks_global,
                                  d.z_chunk0, ax.band, d.ny * d.nx, n_lines, n_tiles, peers, (long long)(timeout_s * 1.9e9), status);
  HS2_CUDA_CHECK(cudaGetLastError());
  pl->last_kernel[2] = HS2_K_Z_SLAB;
  return HS2_OK;
}

template <int M, int W>
int launch_zfused_w(hs2_plan *pl, double *data, const double *Tin, double *Tout, double *Yall, const ZPeers &peers,
                    double timeout_s, int *status, int *tile_lines, cudaStream_t st) {
  const hs2_plan_desc &d = pl->d;
  const hs2_axis_tables &ax = d.axis[2];
  const int P_loc = (int)(d.nz / M);
  const int G = 256 / (W * P_loc) > 0 ? 256 / (W * P_loc) : 1;
  if (tile_lines) {
    *tile_lines = W * G;
    return HS2_OK;
  }
  const int n_lines = (int)(d.ny * d.nx);
  const int n_tiles = (n_lines + W - 1) / W;
  const int row0 = d.z_chunk0 * M;
  const int nz_loc = (int)d.nz;
  const size_t smem = ((size_t)HS2_T_PLANES * nz_loc + (size_t)(P_loc + 1) * 2 * d.z_chunks_global + (size_t)M * 256) * sizeof(double);
  HS2_REQUIRE(smem <= 110 * 1024, "fused z sweep: tables and the parked tile need %zu B of shared memory", smem);
  auto kern = z_fused<M, W>;
  HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = pl->sm_count * 2;
  if (blocks > (n_tiles + G - 1) / G) blocks = (n_tiles + G - 1) / G;
  dim3 block(W, P_loc, G);
  kern<<<blocks, block, smem, st>>>(data, Tin, Tout, ax.d_line_id, ax.d_tab, ax.d_GE, Yall, ax.pitch, row0, nz_loc, d.z_chunks_global,
                                    d.z_chunk0, ax.band, d.ny * d.nx, n_lines, n_tiles, peers, (long long)(timeout_s * 1.9e9),
                                    status);
  HS2_CUDA_CHECK(cudaGetLastError());
  pl->last_kernel[2] = HS2_K_Z_SLAB;
  return HS2_OK;
}

template <int M>
int launch_zfused(hs2_plan *pl, double *data, const double *Tin, double *Tout, double *Yall, const ZPeers &peers, double timeout_s,
                  int *status, int *tile_lines, cudaStream_t st) {
  const int P_loc = (int)(pl->d.nz / M);
  HS2_REQUIRE(P_loc * 8 <= 256, "distributed z sweep: %d local chunks of %d rows do not fit a block", P_loc, M);
  // thin slabs: all chunks of a line in one warp (8 lines x 4 chunks, 16 x 2, 32 x 1), no barriers
  static const bool no_warp = getenv("HS2_ZFUSED_BLOCK") != nullptr && getenv("HS2_ZFUSED_BLOCK")[0] == '1';
  if (!no_warp && pl->d.nz % M == 0) {
    if (P_loc == 4) return launch_zfused_warp<M, 8>(pl, data, Tin, Tout, Yall, peers, timeout_s, status, tile_lines, st);
    if (P_loc == 2) return launch_zfused_warp<M, 16>(pl, data, Tin, Tout, Yall, peers, timeout_s, status, tile_lines, st);
    if (P_loc == 1) return launch_zfused_warp<M, 32>(pl, data, Tin, Tout, Yall, peers, timeout_s, status, tile_lines, st);
  }
  if (P_loc * 16 <= 256) return launch_zfused_w<M, 16>(pl, data, Tin, Tout, Yall, peers, timeout_s, status, tile_lines, st);
  return launch_zfused_w<M, 8>(pl, data, Tin, Tout, Yall, peers, timeout_s, status, tile_lines, st);
}

template <int M, int W>
int launch_zdist_w(hs2_plan *pl, int phase, double *data, const double *Tin, double *Tout, double *Y, int line0,
                   int n_lines, const ZPeers &peers, bool full_cols, cudaStream_t st) {
  const hs2_plan_desc &d = pl->d;
  const hs2_axis_tables &ax = d.axis[2];
  const int P_loc = (int)(d.nz / M);
  const int G = 256 / (W * P_loc) > 0 ? 256 / (W * P_loc) : 1;
  const int n_tiles = (n_lines + W - 1) / W;
  const int row0 = d.z_chunk0 * M;
  const int nz_loc = (int)d.nz;
  const size_t smem = ((size_t)HS2_T_PLANES * nz_loc + (phase ? (size_t)(P_loc + 1) * 2 * d.z_chunks_global : 0)) * sizeof(double);
  HS2_REQUIRE(smem <= 100 * 1024, "distributed z sweep: tables need %zu B of shared memory", smem);
  int blocks = pl->sm_count * 2;
  if (blocks > (n_tiles + G - 1) / G) blocks = (n_tiles + G - 1) / G;
  dim3 block(W, P_loc, G);
  // interface rows: one column per line of this launch, or (full_cols) per line of the slab
  const int64_t ldy = full_cols ? d.ny * d.nx : n_lines;
  const int ycol0 = full_cols ? line0 : 0;
  if (phase == 0) {
    auto kern = z_forward<M, W>;
    if (smem > 48 * 1024) HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, block, smem, st>>>(data, ax.d_line_id, ax.d_tab, Y, ax.pitch, row0, nz_loc, d.ny * d.nx, line0, n_lines,
                                      n_tiles, peers, ldy, ycol0);
  } else {
    auto kern = z_backward<M, W>;
    if (smem > 48 * 1024) HS2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, block, smem, st>>>(data, Tin, Tout, ax.d_line_id, ax.d_tab, ax.d_GE, Y, ax.pitch, row0, nz_loc,
                                      d.z_chunks_global, d.z_chunk0, ax.band, d.ny * d.nx, line0, n_lines, n_tiles, ldy, ycol0);
  }
  HS2_CUDA_CHECK(cudaGetLastError());
  return HS2_OK;
}

template <int M>
int launch_zdist(hs2_plan *pl, int phase, double *data, const double *Tin, double *Tout, double *Y, int line0,
                 int n_lines, const ZPeers &peers, bool full_cols, cudaStream_t st) {
  const int P_loc = (int)(pl->d.nz / M);
  HS2_REQUIRE(P_loc * 8 <= 256, "distributed z sweep: %d local chunks of %d rows do not fit a block", P_loc, M);
  if (P_loc * 16 <= 256) return launch_zdist_w<M, 16>(pl, phase, data, Tin, Tout, Y, line0, n_lines, peers, full_cols, st);
  return launch_zdist_w<M, 8>(pl, phase, data, Tin, Tout, Y, line0, n_lines, peers, full_cols, st);
}

int hs2_zdist(hs2_plan *pl, int phase, double *data, const double *Tin, double *Tout, double *Y, int64_t line0,
              int64_t n_lines, int n_peers, const uint64_t *peer_y, bool full_cols, cudaStream_t st) {
  const hs2_plan_desc &d = pl->d;
  const int M = d.axis[2].chunk;
  HS2_REQUIRE(d.z_chunks_global > 0, "plan is not part of a z-slab decomposition");
  HS2_REQUIRE((M == 8 || M == 16 || M == 32) && d.nz % M == 0 && d.axis[2].d_tab && d.axis[2].d_GE,
              "distributed z sweep needs chunk tables and nz (%lld) divisible by the chunk size (%d)", (long long)d.nz, M);
  HS2_REQUIRE(d.ny * d.nx < ((int64_t)1 << 31), "grid too large");
  HS2_REQUIRE(line0 >= 0 && n_lines > 0 && line0 + n_lines <= d.ny * d.nx, "distributed z sweep: bad line range");
  HS2_REQUIRE(n_peers >= 0 && n_peers <= HS2_MAX_Z_PEERS && (n_peers == 0 || peer_y), "distributed z sweep: %d peers (max %d)",
              n_peers, HS2_MAX_Z_PEERS);
  pl->last_kernel[2] = HS2_K_Z_SLAB;
  ZPeers peers;
  peers.n = n_peers;
  for (int q = 0; q < HS2_MAX_Z_PEERS; ++q) peers.y[q] = q < n_peers ? reinterpret_cast<double *>(peer_y[q]) : nullptr;
  switch (M) {
    case 8: return launch_zdist<8>(pl, phase, data, Tin, Tout, Y, (int)line0, (int)n_lines, peers, full_cols, st);
    case 16: return launch_zdist<16>(pl, phase, data, Tin, Tout, Y, (int)line0, (int)n_lines, peers, full_cols, st);
    default: return launch_zdist<32>(pl, phase, data, Tin, Tout, Y, (int)line0, (int)n_lines, peers, full_cols, st);
  }
}

// fused slab z sweep; tile_lines != NULL: only report the lines per tile
int hs2_zfused(hs2_plan *pl, double *data, const double *Tin, double *Tout, double *Yall, int n_peers, const uint64_t *peer_y,
               double timeout_s, int *status, int *tile_lines, cudaStream_t st) {
  const hs2_plan_desc &d = pl->d;
  const int M = d.axis[2].chunk;
  HS2_REQUIRE(d.z_chunks_global > 0, "plan is not part of a z-slab decomposition");
  HS2_REQUIRE((M == 8 || M == 16 || M == 32) && d.nz % M == 0 && d.axis[2].d_tab && d.axis[2].d_GE,
              "distributed z sweep needs chunk tables and nz (%lld) divisible by the chunk size (%d)", (long long)d.nz, M);
  HS2_REQUIRE(d.ny * d.nx < ((int64_t)1 << 31), "grid too large");
  ZPeers peers;
  peers.n = 0;
  if (!tile_lines) {
    HS2_REQUIRE(n_peers >= 0 && n_peers <= HS2_MAX_Z_PEERS && (n_peers == 0 || peer_y) && status, "fused z sweep: %d peers (max %d)",
                n_peers, HS2_MAX_Z_PEERS);
    peers.n = n_peers;
    for (int q = 0; q < HS2_MAX_Z_PEERS; ++q) peers.y[q] = q < n_peers ? reinterpret_cast<double *>(peer_y[q]) : nullptr;
  }
  switch (M) {
    case 8: return launch_zfused<8>(pl, data, Tin, Tout, Yall, peers, timeout_s, status, tile_lines, st);
    case 16: return launch_zfused<16>(pl, data, Tin, Tout, Yall, peers, timeout_s, status, tile_lines, st);
    default: return launch_zfused<32>(pl, data, Tin, Tout, Yall, peers, timeout_s, status, tile_lines, st);
  }
}

int hs2_fill_empty(void *d_ptr, int64_t n_doubles, cudaStream_t st) {
  if (n_doubles <= 0) return HS2_OK;
  fill_u64_kernel<<<256, 256, 0, st>>>(reinterpret_cast<unsigned long long *>(d_ptr), n_doubles, HS2_Y_EMPTY);
  HS2_CUDA_CHECK(cudaGetLastError());
  return HS2_OK;
}

bool hs2_tile_supported(const hs2_plan *p, int axis) {
  const hs2_axis_tables &ax = p->d.axis[axis];
  if (axis == 2 && p->d.z_chunks_global > 0) return false;   // slab plans use hs2_sweep_z_forward/backward
  if (p->d.flags & HS2_FLAG_FORCE_FALLBACK) return false;
  if (!(ax.chunk == 8 || ax.chunk == 16 || ax.chunk == 32)) return false;
  if (!ax.d_tab || !ax.d_GE || ax.pitch <= 0 || (ax.pitch & 1)) return false;
  const int maxthreads = ax.chunk >= 32 ? 256 : 512;
  return ax.n_chunks >= 1 && ax.n_chunks * 8 <= maxthreads;
}

int hs2_tile_sweep_y(hs2_plan *p, double *W, cudaStream_t st) {
  const hs2_plan_desc &d = p->d;
  HS2_REQUIRE(d.nx < ((int64_t)1 << 31) && d.nz < ((int64_t)1 << 31), "grid too large");
  return dispatch<false>(p, d.axis[1], W, nullptr, nullptr, (int)d.ny, d.nx, (int)d.nz, (int)d.nx, d.ny * d.nx, st);
}

int hs2_tile_sweep_z(hs2_plan *p, const double *T, double *Tout, double *W, cudaStream_t st) {
  const hs2_plan_desc &d = p->d;
  HS2_REQUIRE(d.ny * d.nx < ((int64_t)1 << 31), "grid too large");
  return dispatch<true>(p, d.axis[2], W, T, Tout, (int)d.nz, d.ny * d.nx, 1, (int)(d.ny * d.nx), 0, st);
}

#ifdef HS2_PHASE_TIMING
extern "C" int hs2_debug_phase_strided(unsigned long long *out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_hs2_phase, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_hs2_phase, z, sizeof(z));
  }
  return 0;
}
#endif

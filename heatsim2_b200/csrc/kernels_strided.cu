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

namespace {

template <int M, int W, bool FINAL>
__global__ void __launch_bounds__(M >= 32 ? 256 : 512, 2)
strided_sweep(double *__restrict__ data, const double *__restrict__ Tin, double *__restrict__ Tout,
              const uint32_t *__restrict__ line_id, const double *__restrict__ tab, const double *__restrict__ GE,
              int L, int pitch, int P,
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
  const double E = chunk_interface(ge, Y, P, W, w);
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
                                                                      ax.pitch, P, stride, tiles_per_group,
                                                                      lines_per_group, group_stride);
  else
    strided_sweep<M, 8, FINAL><<<(unsigned)blocks, block, smem, st>>>(data, Tin, Tout, ax.d_line_id, ax.d_tab, ax.d_GE, L,
                                                                     ax.pitch, P, stride, tiles_per_group,
                                                                     lines_per_group, group_stride);
  HS2_CUDA_CHECK(cudaGetLastError());
  return HS2_OK;
}

template <bool FINAL>
int dispatch(const hs2_axis_tables &ax, double *data, const double *Tin, double *Tout, int L, int64_t stride,
             int n_groups, int lines_per_group, int64_t group_stride, cudaStream_t st) {
  switch (ax.chunk) {
    case 8: return launch<8, FINAL>(ax, data, Tin, Tout, L, stride, n_groups, lines_per_group, group_stride, st);
    case 16: return launch<16, FINAL>(ax, data, Tin, Tout, L, stride, n_groups, lines_per_group, group_stride, st);
    case 32: return launch<32, FINAL>(ax, data, Tin, Tout, L, stride, n_groups, lines_per_group, group_stride, st);
  }
  hs2_set_error("strided sweep: unsupported chunk size %d", ax.chunk);
  return HS2_E_INVALID;
}

}  // namespace

bool hs2_tile_supported(const hs2_plan *p, int axis) {
  const hs2_axis_tables &ax = p->d.axis[axis];
  if (p->d.flags & HS2_FLAG_FORCE_FALLBACK) return false;
  if (!(ax.chunk == 8 || ax.chunk == 16 || ax.chunk == 32)) return false;
  if (!ax.d_tab || !ax.d_GE || ax.pitch <= 0 || (ax.pitch & 1)) return false;
  const int maxthreads = ax.chunk >= 32 ? 256 : 512;
  return ax.n_chunks >= 1 && ax.n_chunks * 8 <= maxthreads;
}

int hs2_tile_sweep_y(hs2_plan *p, double *W, cudaStream_t st) {
  const hs2_plan_desc &d = p->d;
  HS2_REQUIRE(d.nx < ((int64_t)1 << 31) && d.nz < ((int64_t)1 << 31), "grid too large");
  return dispatch<false>(d.axis[1], W, nullptr, nullptr, (int)d.ny, d.nx, (int)d.nz, (int)d.nx, d.ny * d.nx, st);
}

int hs2_tile_sweep_z(hs2_plan *p, const double *T, double *Tout, double *W, cudaStream_t st) {
  const hs2_plan_desc &d = p->d;
  HS2_REQUIRE(d.ny * d.nx < ((int64_t)1 << 31), "grid too large");
  return dispatch<true>(d.axis[2], W, T, Tout, (int)d.nz, d.ny * d.nx, 1, (int)(d.ny * d.nx), 0, st);
}

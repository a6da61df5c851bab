// x_common.cuh - pieces shared by the two x-sweep kernels (kernels_x.cu: one
// tile per step; kernels_xm.cu: blocks that march along z and re-use planes).
#pragma once
#include "chunk_core.cuh"

constexpr int HS2_XR = 8;  // x-lines per block tile

struct SrcTab {  // active volumetric classes of this step (<= 8)
  int n;
  uint8_t idx[8];
  double val[8];
};

// row pitch (doubles) of the chunk-padded shared-memory tile: cell (r, i) lives
// at r*Sr + i + i/M; Sr = 2 (mod 16) makes the chunk-major accesses of phase 2
// (lanes = 8 lines x 4 chunks) conflict free for 8-byte words
__host__ __device__ inline int hs2_row_pitch(int P, int M) {
  int s = P * (M + 1);
  while ((s & 15) != 2) ++s;
  return s;
}

inline int hs2_make_src_tab(const hs2_source *src, SrcTab *st) {
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

// Phases 2 and 3 of an x tile: the right hand sides of HS2_XR lines sit in
// `buf` (chunk-padded); thread (r = tid % R, p = tid / R) solves chunk p of
// line r in registers, the solution goes back through `buf` and is stored
// with 16-byte coalesced writes.  Must be called by all R*P threads; ends with
// the data in global memory but WITHOUT a trailing barrier.
template <int M>
__device__ __forceinline__ void hs2_x_solve_store(double *buf, double *Y, double *Es, int Sr, int tid, int nthreads,
                                                  int nrows, int nx, int P, int band, int pitch, uint32_t lid,
                                                  uint32_t lid_c, const double *s_tab, const double *s_ge,
                                                  const double *__restrict__ tab, const double *__restrict__ GE,
                                                  double *__restrict__ Wrow0) {
  constexpr int R = HS2_XR;
  const int r = tid % R;
  const int p = tid / R;
  const bool live = r < nrows;
  const int c0 = p * M;
  const int rows = min(M, nx - c0);
  const bool full = rows == M;
  const double *tb = lid == lid_c ? s_tab + c0 : tab + ((int64_t)lid * HS2_T_PLANES) * pitch + c0;
  const double *ge = lid == lid_c ? s_ge + p * (2 * P) : GE + ((int64_t)lid * P + p) * (2 * P);
  double *mine = buf + r * Sr + p * (M + 1);
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
  Y[(2 * p) * R + r] = yf;
  Y[(2 * p + 1) * R + r] = last;
  __syncthreads();
  const double E = chunk_interface(ge, Y, P, R, r, p, band);
  Es[p * R + r] = E;
  __syncthreads();
  const double alpha = p > 0 ? Es[(p - 1) * R + r] : 0.0;
  if (full)
    chunk_backward_full<M>(v, tb, pitch, alpha, E);
  else
    chunk_backward_short<M>(v, tb, pitch, rows, alpha, E);
  if (live) {
#pragma unroll
    for (int t = 0; t < M; ++t)
      if (t < rows) mine[t] = v[t];
  }
  __syncthreads();
  for (int i = 2 * tid; i < nx; i += 2 * nthreads) {
    const double *bcol0 = buf + i + i / M;
    const double *bcol1 = buf + (i + 1) + (i + 1) / M;
    double2 o[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) o[rr] = make_double2(bcol0[rr * Sr], bcol1[rr * Sr]);
#pragma unroll
    for (int rr = 0; rr < R; ++rr)
      if (rr < nrows) *reinterpret_cast<double2 *>(Wrow0 + (int64_t)rr * nx + i) = o[rr];
  }
}

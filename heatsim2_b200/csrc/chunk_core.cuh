// chunk_core.cuh - per-thread arithmetic of the partitioned tridiagonal solve,
// shared by the strided (y, z) and the contiguous (x) sweep kernels.
//
// A thread owns M consecutive rows of one line in registers.  Table pointers
// may point to global memory or to a block's shared-memory copy (plain loads).  Tables (one set
// per unique line, planes of `pitch` doubles, see heatsim2_b200/plan.py
// chunk_factors): INV, F, C for the forward part, S, CP for the backward part.
// `tb` points at the chunk's first row inside plane 0.
#pragma once
#include "hs2_common.cuh"

// forward elimination of a full chunk; returns the first-row functional yf
// and leaves the chunk's local forward solution in v (v[M-1] is y_last)
template <int M>
__device__ __forceinline__ double chunk_forward_full(double (&v)[M], const double *__restrict__ tb, int pitch) {
  {
    const double2 *ci = reinterpret_cast<const double2 *>(tb + HS2_T_INV * pitch);
#pragma unroll
    for (int t = 0; t < M; t += 2) {
      const double2 c = ci[t / 2];
      v[t] *= c.x;
      v[t + 1] *= c.y;
    }
  }
  {
    const double2 *cf = reinterpret_cast<const double2 *>(tb + HS2_T_F * pitch);
    double prev = 0.0;
#pragma unroll
    for (int t = 0; t < M; t += 2) {
      const double2 c = cf[t / 2];
      prev = fma(-c.x, prev, v[t]);
      v[t] = prev;
      prev = fma(-c.y, prev, v[t + 1]);
      v[t + 1] = prev;
    }
  }
  const double2 *cc = reinterpret_cast<const double2 *>(tb + HS2_T_C * pitch);
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int t = 0; t < M; t += 2) {
    const double2 c = cc[t / 2];
    a0 = fma(c.x, v[t], a0);
    a1 = fma(c.y, v[t + 1], a1);
  }
  return a0 + a1;
}

// same for a chunk of `rows` < M rows (last chunk of a line); *last = y_last
template <int M>
__device__ __forceinline__ double chunk_forward_short(double (&v)[M], const double *__restrict__ tb, int pitch, int rows,
                                                      double *last) {
  double prev = 0.0, yf = 0.0;
#pragma unroll
  for (int t = 0; t < M; ++t) {
    if (t < rows) {
      prev = fma(-tb[HS2_T_F * pitch + t], prev, v[t] * tb[HS2_T_INV * pitch + t]);
      v[t] = prev;
      yf = fma(tb[HS2_T_C * pitch + t], prev, yf);
    }
  }
  *last = prev;
  return yf;
}

// back substitution given alpha (true x just before the chunk) and E (true x
// of the chunk's last row); v becomes the solution
template <int M>
__device__ __forceinline__ void chunk_backward_full(double (&v)[M], const double *__restrict__ tb, int pitch, double alpha,
                                                    double E) {
  {
    const double2 *cs = reinterpret_cast<const double2 *>(tb + HS2_T_S * pitch);
#pragma unroll
    for (int t = 0; t < M; t += 2) {
      const double2 c = cs[t / 2];
      v[t] = fma(-alpha, c.x, v[t]);
      v[t + 1] = fma(-alpha, c.y, v[t + 1]);
    }
  }
  const double2 *cp = reinterpret_cast<const double2 *>(tb + HS2_T_CP * pitch);
  double nxt = E;
  v[M - 1] = E;
#pragma unroll
  for (int t = M - 2; t >= 0; t -= 2) {
    const double2 c = cp[t / 2];  // (cp[t], cp[t+1])
    if (t + 1 < M - 1) {
      nxt = fma(-c.y, nxt, v[t + 1]);
      v[t + 1] = nxt;
    }
    nxt = fma(-c.x, nxt, v[t]);
    v[t] = nxt;
  }
}

template <int M>
__device__ __forceinline__ void chunk_backward_short(double (&v)[M], const double *__restrict__ tb, int pitch, int rows,
                                                     double alpha, double E) {
  double nxt = E;
#pragma unroll
  for (int t = M - 1; t >= 0; --t) {
    if (t < rows) {
      if (t < rows - 1)
        nxt = fma(-tb[HS2_T_CP * pitch + t], nxt, fma(-alpha, tb[HS2_T_S * pitch + t], v[t]));
      v[t] = nxt;
    }
  }
}

// ---------------------------------------------------------------------------
// Uniform-chunk fast path.  The chunk-local factorisation restarts at every
// chunk start, so all interior chunks of a line with constant coefficients have
// bit-identical factor tables (and so do the interior chunks of every other
// line through the same material).  The host finds the most common chunk table
// of an axis (plan.py uniform_chunks), the kernels receive it BY VALUE as a
// __grid_constant__ parameter - i.e. in the constant bank - and a warp whose
// chunks all carry it reads every factor as a constant operand of the DFMA
// itself: no load instruction, no L1/shared-memory wavefront.  Measured in
// round 1 (profiles/kernels_r01e.md): the broadcast table loads were 85 % of
// the L1 data-pipe wavefronts of the y sweep and a third of the x sweep's.
// Same values, same operation order => bit-identical to the table-load path.
// (struct UTab: hs2_common.cuh)
struct NoPace {
  __device__ __forceinline__ void one() {}
};
// PACE: something with one() that is called once per two rows of the dependent chain (the persistent z kernel
// issues the next tile's cp.async pieces there: the chain leaves the issue slots free)
template <int M, class PACE>
__device__ __forceinline__ double chunk_forward_const(double (&v)[M], const UTab &ut, PACE &pace) {
#pragma unroll
  for (int t = 0; t < M; ++t) v[t] *= ut.v[HS2_T_INV][t];
  double prev = 0.0;
#pragma unroll
  for (int t = 0; t < M; ++t) {
    prev = fma(-ut.v[HS2_T_F][t], prev, v[t]);
    v[t] = prev;
    if (t & 1) pace.one();
  }
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int t = 0; t < M; t += 2) {
    a0 = fma(ut.v[HS2_T_C][t], v[t], a0);
    a1 = fma(ut.v[HS2_T_C][t + 1], v[t + 1], a1);
  }
  return a0 + a1;
}

template <int M>
__device__ __forceinline__ double chunk_forward_const(double (&v)[M], const UTab &ut) {
  NoPace np;
  return chunk_forward_const<M, NoPace>(v, ut, np);
}

template <int M>
__device__ __forceinline__ void chunk_backward_const(double (&v)[M], const UTab &ut, double alpha, double E) {
#pragma unroll
  for (int t = 0; t < M; ++t) v[t] = fma(-alpha, ut.v[HS2_T_S][t], v[t]);
  double nxt = E;
  v[M - 1] = E;
#pragma unroll
  for (int t = M - 2; t >= 0; --t) {
    nxt = fma(-ut.v[HS2_T_CP][t], nxt, v[t]);
    v[t] = nxt;
  }
}

// E_p = row p of the inverse interface operator applied to the interleaved
// (yf_0, yl_0, yf_1, yl_1, ...) values of this line held in Y[2P][ld] column w.
// The operator decays geometrically away from the diagonal; `band` (from the
// host, chunk_factors) is the half-width beyond which every entry is below
// 1e-16 of the row maximum (plan.py interface_band), so only chunks p-band..p+band are visited.
__device__ __forceinline__ double chunk_interface(const double *__restrict__ ge, const double *Y, int P, int ld, int w,
                                                  int p, int band) {
  double e0 = 0.0, e1 = 0.0;
  const double2 *g2 = reinterpret_cast<const double2 *>(ge);
  const int q0 = max(0, p - band), q1 = min(P - 1, p + band);
#pragma unroll 4
  for (int q = q0; q <= q1; ++q) {
    const double2 g = g2[q];
    e0 = fma(g.x, Y[(2 * q) * ld + w], e0);
    e1 = fma(g.y, Y[(2 * q + 1) * ld + w], e1);
  }
  return e0 + e1;
}

// ---------------------------------------------------------------------------
// Accessor-based variants.  Measured on B200 (profiles/micro/tabsrc.cu): the
// per-row factors are broadcast loads (all lines of a warp that sit in the same
// chunk read the same address), and a 16-byte shared load costs the data pipe
// ~8 cycles per warp whatever the addresses are, while 8-byte shared loads of
// one or two distinct addresses cost 1-2.  The solve went from 1.5 to 4.3-7.0
// row-updates/cycle/SM.  So the tables are read with explicit 8-byte
// ld.shared.f64 (inline PTX keeps the compiler from re-fusing neighbours into
// one 16-byte load) when the block has them in shared memory, and with scalar
// __ldg otherwise.
struct TabShared {
  uint32_t a;        // shared-window byte address of plane 0 at this thread's first row
  uint32_t pitch_b;  // bytes between planes
  __device__ __forceinline__ uint32_t plane(int pl) const { return a + pl * pitch_b; }
  __device__ __forceinline__ static double ld(uint32_t q, int t) {
    double x;
    asm("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(q + 8u * (uint32_t)t));
    return x;
  }
};

struct TabGlobal {
  const double *b;  // plane 0 at this thread's first row
  int pitch;        // doubles between planes
  __device__ __forceinline__ const double *plane(int pl) const { return b + (int64_t)pl * pitch; }
  __device__ __forceinline__ static double ld(const double *q, int t) { return __ldg(q + t); }
};

// forward elimination of a chunk of `rows` <= M rows (rows == M: no guards);
// returns y_first, *last = y_last, v = chunk-local forward solution
template <int M, bool FULL, class TAB>
__device__ __forceinline__ double chunk_fwd(double (&v)[M], const TAB &tb, int rows, double *last) {
  {
    const auto q = tb.plane(HS2_T_INV);
#pragma unroll
    for (int t = 0; t < M; ++t)
      if (FULL || t < rows) v[t] *= TAB::ld(q, t);
  }
  double prev = 0.0;
  {
    const auto q = tb.plane(HS2_T_F);
#pragma unroll
    for (int t = 0; t < M; ++t) {
      if (FULL || t < rows) {
        prev = fma(-TAB::ld(q, t), prev, v[t]);
        v[t] = prev;
      }
    }
  }
  *last = prev;
  double a0 = 0.0, a1 = 0.0;
  {
    const auto q = tb.plane(HS2_T_C);
#pragma unroll
    for (int t = 0; t < M; t += 2) {
      if (FULL || t < rows) a0 = fma(TAB::ld(q, t), v[t], a0);
      if (FULL || t + 1 < rows) a1 = fma(TAB::ld(q, t + 1), v[t + 1], a1);
    }
  }
  return a0 + a1;
}

// back substitution given alpha (true x just before the chunk) and E (true x of
// the chunk's last row); v becomes the solution
template <int M, bool FULL, class TAB>
__device__ __forceinline__ void chunk_bwd(double (&v)[M], const TAB &tb, int rows, double alpha, double E) {
  {
    const auto q = tb.plane(HS2_T_S);
#pragma unroll
    for (int t = 0; t < M; ++t)
      if (FULL || t < rows) v[t] = fma(-alpha, TAB::ld(q, t), v[t]);
  }
  const auto q = tb.plane(HS2_T_CP);
  double nxt = E;
#pragma unroll
  for (int t = M - 1; t >= 0; --t) {
    if (FULL) {
      if (t < M - 1) nxt = fma(-TAB::ld(q, t), nxt, v[t]);
      v[t] = nxt;
    } else if (t < rows) {
      if (t < rows - 1) nxt = fma(-TAB::ld(q, t), nxt, v[t]);
      v[t] = nxt;
    }
  }
}

// tridiag.cu - device drop-ins for heatsim2/tridiag.pyx.
//
// The reference factors and solves ONE tridiagonal chain of n = nz*ny*nx rows
// on one CPU thread.  Each of its three loops is a first-order recurrence:
//   LU       piv[r+1] = b[r+1] - a[r+1]*c[r]/piv[r]          (tridiag.pyx:25-41)
//   forward  x[r]     = (rhs[r] - L[r,0]*x[r-1]) / L[r,1]    (tridiag.pyx:58-61)
//   backward x[r]     = x[r] - U[r,2]*x[r+1]                 (tridiag.pyx:65-67)
// A recurrence x -> f_r(x) with f_r affine (solve) or Moebius (LU) is a scan
// over function composition, so the chain is cut into tiles: pass 1 composes
// each tile into one map, pass 2 scans the tile maps, pass 3 re-runs the
// reference's own arithmetic inside every tile from the now-known entry value.
// Only the tile entry values differ in rounding from a serial evaluation.
#include "hs2_common.cuh"

namespace {

constexpr int TPB = 256;       // threads per block
constexpr int ITEMS = 8;       // rows per thread
constexpr int TILE = TPB * ITEMS;

// ---------------------------------------------------------------- monoids
struct Affine {  // x -> a*x + b
  double a, b;
  __device__ static Affine identity() { return {1.0, 0.0}; }
  __device__ double eval(double x) const { return fma(a, x, b); }
};
__device__ inline Affine then(const Affine &f, const Affine &g) {  // g after f
  return {g.a * f.a, fma(g.a, f.b, g.b)};
}
__device__ inline Affine shfl_up(const Affine &v, int d) {
  return {__shfl_up_sync(0xffffffffu, v.a, d), __shfl_up_sync(0xffffffffu, v.b, d)};
}

struct Moebius {  // p -> (m00*p + m01) / (m10*p + m11)
  double m00, m01, m10, m11;
  __device__ static Moebius identity() { return {1.0, 0.0, 0.0, 1.0}; }
  __device__ double eval(double p) const { return fma(m00, p, m01) / fma(m10, p, m11); }
};
__device__ inline Moebius then(const Moebius &f, const Moebius &g) {
  Moebius r{fma(g.m00, f.m00, g.m01 * f.m10), fma(g.m00, f.m01, g.m01 * f.m11),
            fma(g.m10, f.m00, g.m11 * f.m10), fma(g.m10, f.m01, g.m11 * f.m11)};
  // projective: rescale so that products over long chains stay in range
  const double s = fmax(fmax(fabs(r.m00), fabs(r.m01)), fmax(fabs(r.m10), fabs(r.m11)));
  if (s > 0.0) {
    const double inv = 1.0 / s;
    r.m00 *= inv; r.m01 *= inv; r.m10 *= inv; r.m11 *= inv;
  }
  return r;
}
__device__ inline Moebius shfl_up(const Moebius &v, int d) {
  return {__shfl_up_sync(0xffffffffu, v.m00, d), __shfl_up_sync(0xffffffffu, v.m01, d),
          __shfl_up_sync(0xffffffffu, v.m10, d), __shfl_up_sync(0xffffffffu, v.m11, d)};
}

// --------------------------------------------------------- recurrence kinds
// Each "problem" exposes: Map (monoid), map(r) = the r-th step as a Map,
// start value before row 0, and emit(r, x_prev) -> x_r doing the reference's
// own arithmetic and writing the outputs.
struct FwdSolve {
  using Map = Affine;
  const double *L, *b;
  double *x;
  int64_t n;
  __device__ Map map(int64_t s) const {
    const double piv = L[3 * s + 1];
    return {s == 0 ? 0.0 : -L[3 * s] / piv, b[s] / piv};
  }
  __device__ double emit(int64_t s, double prev) const {
    const double v = s == 0 ? b[0] / L[1] : (b[s] - L[3 * s] * prev) / L[3 * s + 1];
    x[s] = v;
    return v;
  }
};

struct BwdSolve {  // sequence index s runs from the last row to the first
  using Map = Affine;
  const double *U;
  double *x;
  int64_t n;
  __device__ int64_t row(int64_t s) const { return n - 1 - s; }
  __device__ Map map(int64_t s) const {
    const int64_t r = row(s);
    return {s == 0 ? 0.0 : -U[3 * r + 2], x[r]};
  }
  __device__ double emit(int64_t s, double prev) const {
    const int64_t r = row(s);
    const double v = s == 0 ? x[r] : x[r] - U[3 * r + 2] * prev;
    x[r] = v;
    return v;
  }
};

struct LuPivots {  // state = pivot of the row
  using Map = Moebius;
  const double *A;
  double *Lm, *Um;
  int64_t n;
  __device__ Map map(int64_t s) const {
    if (s == 0) return {0.0, A[1], 0.0, 1.0};              // constant map -> A[0,1]
    // piv_s = b_s - a_s * c_{s-1} / piv_{s-1}
    return {A[3 * s + 1], -A[3 * s] * A[3 * (s - 1) + 2], 1.0, 0.0};
  }
  __device__ double emit(int64_t s, double prev) const {
    double piv, sub;
    if (s == 0) {
      piv = A[1];
      sub = 0.0;
    } else {
      sub = A[3 * s];
      const double u2prev = A[3 * (s - 1) + 2] / prev;     // Umat[s-1,2]
      piv = A[3 * s + 1] - u2prev * sub;
    }
    Lm[3 * s] = sub;
    Lm[3 * s + 1] = piv;
    Lm[3 * s + 2] = 0.0;
    Um[3 * s] = 0.0;
    Um[3 * s + 1] = 1.0;
    Um[3 * s + 2] = A[3 * s + 2] / piv;
    return piv;
  }
};

// ------------------------------------------------------------------ passes
template <class P>
__device__ inline typename P::Map thread_aggregate(const P &p, int64_t s0, int64_t s1) {
  typename P::Map m = P::Map::identity();
  for (int64_t s = s0; s < s1; ++s) m = then(m, p.map(s));
  return m;
}

// inclusive scan of one Map per thread over the block; returns the exclusive
// prefix of this thread and (to all threads) the block total
template <class Map>
__device__ inline Map block_exclusive(Map v, Map *total) {
  __shared__ Map warp_tot[TPB / 32];
  __shared__ Map blk_tot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  Map inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Map o = shfl_up(inc, d);
    if (lane >= d) inc = then(o, inc);
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    Map t = lane < TPB / 32 ? warp_tot[lane] : Map::identity();
#pragma unroll
    for (int d = 1; d < TPB / 32; d <<= 1) {
      Map o = shfl_up(t, d);
      if (lane >= d) t = then(o, t);
    }
    if (lane < TPB / 32) warp_tot[lane] = t;          // inclusive over warps
    if (lane == TPB / 32 - 1) blk_tot = t;
  }
  __syncthreads();
  Map excl = shfl_up(inc, 1);
  if (lane == 0) excl = Map::identity();
  if (w > 0) excl = then(warp_tot[w - 1], excl);
  *total = blk_tot;
  return excl;
}

template <class P>
__global__ void __launch_bounds__(TPB) pass1(P p, typename P::Map *tile_maps) {
  const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * ITEMS;
  const int64_t s0 = base < p.n ? base : p.n;
  const int64_t s1 = base + ITEMS < p.n ? base + ITEMS : p.n;
  typename P::Map total;
  block_exclusive(thread_aggregate(p, s0, s1), &total);
  if (threadIdx.x == 0) tile_maps[blockIdx.x] = total;
}

// single block: tile_maps[i] <- composition of tiles 0..i-1 (exclusive)
template <class Map>
__global__ void __launch_bounds__(TPB) pass2(Map *tile_maps, int64_t ntiles) {
  __shared__ Map carry;
  if (threadIdx.x == 0) carry = Map::identity();
  __syncthreads();
  for (int64_t t0 = 0; t0 < ntiles; t0 += TPB) {
    const int64_t t = t0 + threadIdx.x;
    Map v = t < ntiles ? tile_maps[t] : Map::identity();
    Map total;
    Map excl = block_exclusive(v, &total);
    const Map c = carry;
    if (t < ntiles) tile_maps[t] = then(c, excl);
    __syncthreads();
    if (threadIdx.x == 0) carry = then(c, total);
    __syncthreads();
  }
}

template <class P>
__global__ void __launch_bounds__(TPB) pass3(P p, const typename P::Map *tile_maps) {
  const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * ITEMS;
  const int64_t s0 = base < p.n ? base : p.n;
  const int64_t s1 = base + ITEMS < p.n ? base + ITEMS : p.n;
  typename P::Map total;
  typename P::Map excl = block_exclusive(thread_aggregate(p, s0, s1), &total);
  // value entering this thread's first row (0 stands for "before row 0"; the
  // first map ignores it)
  double prev = then(tile_maps[blockIdx.x], excl).eval(0.0);
  for (int64_t s = s0; s < s1; ++s) prev = p.emit(s, prev);
}

template <class P>
int run_scan(const P &p, void *scratch, cudaStream_t st) {
  using Map = typename P::Map;
  const int64_t ntiles = (p.n + TILE - 1) / TILE;
  Map *maps = reinterpret_cast<Map *>(scratch);
  pass1<P><<<(unsigned)ntiles, TPB, 0, st>>>(p, maps);
  HS2_CUDA_CHECK(cudaGetLastError());
  pass2<Map><<<1, TPB, 0, st>>>(maps, ntiles);
  HS2_CUDA_CHECK(cudaGetLastError());
  pass3<P><<<(unsigned)ntiles, TPB, 0, st>>>(p, maps);
  HS2_CUDA_CHECK(cudaGetLastError());
  return HS2_OK;
}

}  // namespace

extern "C" {

int64_t hs2_tridiag_scratch_bytes(int64_t n) {
  if (n < 0) n = 0;
  return ((n + TILE - 1) / TILE + 1) * (int64_t)sizeof(Moebius);
}

int hs2_tridiag_lu(int64_t n, const double *d_A, double *d_L, double *d_U, void *d_scratch, void *stream) {
  HS2_REQUIRE(n > 0 && d_A && d_L && d_U && d_scratch, "hs2_tridiag_lu: bad argument");
  HS2_REQUIRE(n < ((int64_t)1 << 31) * (int64_t)TILE, "hs2_tridiag_lu: n too large");
  LuPivots p{d_A, d_L, d_U, n};
  return run_scan(p, d_scratch, (cudaStream_t)stream);
}

int hs2_tridiag_solve(int64_t n, const double *d_L, const double *d_U, const double *d_b, double *d_x, void *d_scratch,
                      void *stream) {
  HS2_REQUIRE(n > 0 && d_L && d_U && d_b && d_x && d_scratch, "hs2_tridiag_solve: bad argument");
  HS2_REQUIRE(d_b != d_x, "hs2_tridiag_solve: d_x must not alias d_b");
  FwdSolve f{d_L, d_b, d_x, n};
  int rc = run_scan(f, d_scratch, (cudaStream_t)stream);
  if (rc) return rc;
  BwdSolve b{d_U, d_x, n};
  return run_scan(b, d_scratch, (cudaStream_t)stream);
}

}  // extern "C"

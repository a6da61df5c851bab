// plan_build.cu - hs2_plan_build: everything between the class tables and a runnable plan, natively.
//
// The reference builds a stage through its C boundary alone: create_adi_step + add_equation per cell
// (heatsim2/alternatingdirection_c.h:50-54, alternatingdirection_c.c:56-199) fill Amat/Bmat/Cmats/Dvec, and
// tridiaglu (heatsim2/tridiag.pyx:9-43) factors A.  Here the same information is the per-cell class id and the
// per-class coefficient row; this file turns them into the tables the kernels read:
//   * unique tridiagonal lines per sweep axis (device: one 64-bit polynomial hash per line, verified exactly
//     against the representative of its group afterwards; host: grouping by first appearance),
//   * Thomas factors of the unique lines (whole-line fallback kernels),
//   * the partitioned-solve tables: chunk-local factorisation, inverse interface operator (dense LU with
//     partial pivoting on the 2P x 2P reduced system), its band, the chunk-interleaved copy, the most common
//     chunk table, the ghost-uniform tables of the warp-per-line x kernel.
// The numpy statement of the same algebra lives in tests/tables_np.py and is what the tests compare with.
#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <map>
#include <new>
#include <thread>
#include <unordered_map>
#include <vector>

#include "hs2_common.cuh"

namespace {

constexpr int XW_HDR = 128;

// ------------------------------------------------------------------------------------------ device side
template <typename CID>
__device__ __forceinline__ int64_t bt_cell(int axis, int64_t line, int r, int64_t ny, int64_t nx) {
  // line numbering: x-lines k*ny+j, y-lines k*nx+i, z-lines j*nx+i
  if (axis == 0) return line * nx + r;
  if (axis == 1) return (line / nx) * (ny * nx) + (int64_t)r * nx + line % nx;
  return (int64_t)r * (ny * nx) + line;
}

// hash[line] = sum_r (sub(line, r) + 1) * pw[r]  (mod 2^64).  x axis: one warp per line; y/z: one thread per
// line (adjacent threads read adjacent cells)
template <typename CID>
__global__ void bt_hash_kernel(const CID *__restrict__ cid, const int32_t *__restrict__ sub, const uint64_t *__restrict__ pw,
                               int axis, int L, int64_t n_lines, int64_t ny, int64_t nx, uint64_t *__restrict__ out) {
  if (axis == 0) {
    const int64_t line = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (line >= n_lines) return;
    uint64_t h = 0;
    for (int r = lane; r < L; r += 32) h += (uint64_t)(sub[cid[line * nx + r]] + 1) * pw[r];
    for (int o = 16; o; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if (lane == 0) out[line] = h;
  } else {
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_lines) return;
    uint64_t h = 0;
    for (int r = 0; r < L; ++r) h += (uint64_t)(sub[cid[bt_cell<CID>(axis, line, r, ny, nx)]] + 1) * pw[r];
    out[line] = h;
  }
}

// reps[u][r] = sub id of row r of line first[u]
template <typename CID>
__global__ void bt_gather_kernel(const CID *__restrict__ cid, const int32_t *__restrict__ sub, const int64_t *__restrict__ first,
                                 int axis, int L, int nu, int64_t ny, int64_t nx, int32_t *__restrict__ reps) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)nu * L) return;
  const int u = (int)(e / L), r = (int)(e % L);
  reps[e] = sub[cid[bt_cell<CID>(axis, first[u], r, ny, nx)]];
}

// every line equals the representative of its group?  (a 64-bit hash collision would show here)
template <typename CID>
__global__ void bt_verify_kernel(const CID *__restrict__ cid, const int32_t *__restrict__ sub, const uint32_t *__restrict__ line_id,
                                 const int32_t *__restrict__ reps, int axis, int L, int64_t n_lines, int64_t ny, int64_t nx,
                                 int *__restrict__ bad) {
  if (axis == 0) {
    const int64_t line = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (line >= n_lines) return;
    const int32_t *rp = reps + (int64_t)line_id[line] * L;
    int diff = 0;
    for (int r = lane; r < L; r += 32) diff |= sub[cid[line * nx + r]] != rp[r];
    if (diff) atomicOr(bad, 1);
  } else {
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_lines) return;
    const int32_t *rp = reps + (int64_t)line_id[line] * L;
    int diff = 0;
    for (int r = 0; r < L; ++r) diff |= sub[cid[bt_cell<CID>(axis, line, r, ny, nx)]] != rp[r];
    if (diff) atomicOr(bad, 1);
  }
}

// a class with a non-zero conductance pointing out of the grid on one of the six faces?
// (alternatingdirection_c.c:160-163 exits there)  out[f] = 1 for face f = z-, z+, y-, y+, x-, x+
template <typename CID>
__global__ void bt_closed_kernel(const CID *__restrict__ cid, const double *__restrict__ coef_raw, int64_t nz, int64_t ny,
                                 int64_t nx, int *__restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nzy = nz * ny, nzx = nz * nx, nyx = ny * nx;
  // columns of the raw coefficient row: M, gx-, gx+, gy-, gy+, gz-, gz+, D
  if (e < nyx) {
    if (coef_raw[(int64_t)cid[e] * 8 + 5] != 0.0) atomicOr(out + 0, 1);
    if (coef_raw[(int64_t)cid[(nz - 1) * nyx + e] * 8 + 6] != 0.0) atomicOr(out + 1, 1);
  }
  if (e < nzx) {
    const int64_t k = e / nx, i = e % nx;
    if (coef_raw[(int64_t)cid[k * nyx + i] * 8 + 3] != 0.0) atomicOr(out + 2, 1);
    if (coef_raw[(int64_t)cid[k * nyx + (ny - 1) * nx + i] * 8 + 4] != 0.0) atomicOr(out + 3, 1);
  }
  if (e < nzy) {
    if (coef_raw[(int64_t)cid[e * nx] * 8 + 1] != 0.0) atomicOr(out + 4, 1);
    if (coef_raw[(int64_t)cid[e * nx + nx - 1] * 8 + 2] != 0.0) atomicOr(out + 5, 1);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------ host algebra
// (exported for the CPU tests through hs2_tables_*; plain C++, no CUDA)

// Thomas factorisation of nu lines [nu][L] -> [nu][L][4] = {1/pivot, lo/pivot, hi/pivot, 0}
// (same recurrence as heatsim2/tridiag.pyx:25-41, stored as reciprocals)
static void bt_thomas(const double *lo, const double *dg, const double *hi, int nu, int L, double *out) {
  for (int u = 0; u < nu; ++u) {
    double cp_prev = 0.0;
    for (int r = 0; r < L; ++r) {
      const int64_t e = (int64_t)u * L + r;
      const double piv = dg[e] - lo[e] * cp_prev;
      const double inv = 1.0 / piv;
      cp_prev = hi[e] * inv;
      out[e * 4 + 0] = inv;
      out[e * 4 + 1] = lo[e] * inv;
      out[e * 4 + 2] = cp_prev;
      out[e * 4 + 3] = 0.0;
    }
  }
}

// in-place inverse of an n x n matrix (row-major) by Gauss-Jordan elimination with partial pivoting
static bool bt_invert(std::vector<double> &A, int n, std::vector<double> &inv) {
  inv.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = std::fabs(A[(size_t)c * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(A[(size_t)r * n + c]) > best) best = std::fabs(A[(size_t)r * n + c]), piv = r;
    if (best == 0.0) return false;
    if (piv != c)
      for (int q = 0; q < n; ++q) {
        std::swap(A[(size_t)piv * n + q], A[(size_t)c * n + q]);
        std::swap(inv[(size_t)piv * n + q], inv[(size_t)c * n + q]);
      }
    const double d = 1.0 / A[(size_t)c * n + c];
    for (int q = 0; q < n; ++q) A[(size_t)c * n + q] *= d, inv[(size_t)c * n + q] *= d;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = A[(size_t)r * n + c];
      if (f == 0.0) continue;
      for (int q = 0; q < n; ++q) A[(size_t)r * n + q] -= f * A[(size_t)c * n + q], inv[(size_t)r * n + q] -= f * inv[(size_t)c * n + q];
    }
  }
  return true;
}

// Partitioned-solve tables of ONE line cut into chunks of M rows (the algebra of tests/tables_np.py chunk_factors):
//   forward   u_k = d_k*inv_k - f_k*u_{k-1} (u_{-1} = 0), y_f = sum_k c_k u_k, y_l = u_last
//   backward  x_k = (u_k - alpha*s_k) - cp_k*x_{k+1}   (x_last = E known, alpha = x before the chunk)
// tab [5][pitch] planes inv, f, c, s, cp; GE [P][2P] rows of the inverse of the interface system that give E_p
// from (y_f0, y_l0, y_f1, ...).  ghost: mirrored neighbours at both ends (x_{-1} = x_0, x_L = x_{L-1}).
static bool bt_chunk_line(const double *lo, const double *dg, const double *hi, int L, int M, int pitch, bool ghost,
                          double *tab, double *GE) {
  const int P = (L + M - 1) / M;
  for (int q = 0; q < 5 * pitch; ++q) tab[q] = 0.0;
  for (int q = 0; q < pitch; ++q) tab[HS2_T_INV * pitch + q] = 1.0;
  std::vector<double> v0(P), cm(P), sl(P), cpl(P);
  for (int p = 0; p < P; ++p) {
    const int r0 = p * M, r1 = std::min(L, (p + 1) * M);
    double cp_prev = 0.0, s_prev = 0.0, c_cur = 1.0, acc = 0.0;
    for (int r = r0; r < r1; ++r) {
      const double piv = dg[r] - (r > r0 ? lo[r] * cp_prev : 0.0);
      const double inv = 1.0 / piv;
      const double f = lo[r] * inv;
      const double cp = hi[r] * inv;
      const double s = r == r0 ? f : -f * s_prev;
      tab[HS2_T_INV * pitch + r] = inv;
      tab[HS2_T_F * pitch + r] = r > r0 ? f : 0.0;    // the kernel's recurrence restarts at a chunk start
      tab[HS2_T_C * pitch + r] = c_cur;
      tab[HS2_T_S * pitch + r] = s;
      tab[HS2_T_CP * pitch + r] = cp;
      acc += c_cur * s;
      c_cur = -cp * c_cur;
      cp_prev = cp, s_prev = s;
    }
    v0[p] = acc, cm[p] = c_cur, sl[p] = s_prev, cpl[p] = cp_prev;
  }
  // reduced system, unknown order [F_0..F_{P-1}, E_0..E_{P-1}]:
  //   F_p + v0_p E_{p-1} - cm_p F_{p+1} = yf_p ;  E_p + sl_p E_{p-1} + cpl_p F_{p+1} = yl_p
  const int n = 2 * P;
  std::vector<double> R((size_t)n * n, 0.0), Ri;
  for (int p = 0; p < P; ++p) {
    R[(size_t)p * n + p] = 1.0;
    R[(size_t)(P + p) * n + P + p] = 1.0;
  }
  for (int p = 0; p < P; ++p) {
    if (p > 0) {
      R[(size_t)p * n + P + p - 1] = v0[p];
      R[(size_t)(P + p) * n + P + p - 1] = sl[p];
    } else if (ghost) {
      R[0] += v0[0];
      R[(size_t)P * n] += sl[0];
    }
    if (p < P - 1) {
      R[(size_t)p * n + p + 1] = -cm[p];
      R[(size_t)(P + p) * n + p + 1] = cpl[p];
    } else if (ghost) {
      R[(size_t)p * n + P + p] += -cm[p];
      R[(size_t)(P + p) * n + P + p] += cpl[p];
    }
  }
  if (!bt_invert(R, n, Ri)) return false;
  for (int p = 0; p < P; ++p)
    for (int q = 0; q < P; ++q) {
      GE[(size_t)p * n + 2 * q] = Ri[(size_t)(P + p) * n + q];
      GE[(size_t)p * n + 2 * q + 1] = Ri[(size_t)(P + p) * n + P + q];
    }
  return true;
}

// half-width, in chunks, outside which every entry of every GE row is below tol times the row maximum
static int bt_band(const double *GE, int nu, int P, double tol) {
  int band = 0;
  for (int u = 0; u < nu; ++u)
    for (int p = 0; p < P; ++p) {
      const double *row = GE + ((size_t)u * P + p) * 2 * P;
      double mx = 0.0;
      for (int q = 0; q < P; ++q) mx = std::max(mx, std::max(std::fabs(row[2 * q]), std::fabs(row[2 * q + 1])));
      for (int q = 0; q < P; ++q)
        if (std::max(std::fabs(row[2 * q]), std::fabs(row[2 * q + 1])) > tol * mx) band = std::max(band, std::abs(q - p));
    }
  return band;
}

struct BtAxis {   // host tables of one axis
  int L = 0, nu = 0, M = 0, P = 0, pitch = 0, band = 0, xw_band = -1;
  std::vector<double> lo, dg, hi;        // [nu][L]
  std::vector<double> lu;                // [nu][L][4]  (or [nu][1][4] when unused)
  std::vector<double> tab, GE, tab_il;   // chunk tables
  std::vector<double> utab;              // [5][M]
  std::vector<uint8_t> ucode;            // [nu][P]
  std::vector<double> xw;                // [nu][XW_HDR + (2 xw_band + 1) * P * 2]
  std::vector<uint8_t> xw_code;          // [nu]
};

static void bt_parallel(int n, const std::function<void(int)> &fn);

// all chunk tables of an axis from its unique rows; weight[u] = lines that use unique line u
static bool bt_axis_tables(BtAxis &a, const std::vector<int64_t> &weight, bool want_lu, bool want_il, bool want_utab,
                           bool want_xw) {
  const int nu = a.nu, L = a.L, M = a.M;
  if (want_lu) {
    a.lu.resize((size_t)nu * L * 4);
    bt_thomas(a.lo.data(), a.dg.data(), a.hi.data(), nu, L, a.lu.data());
  } else {
    a.lu.assign((size_t)nu * 4, 0.0);
  }
  if (M <= 0) return true;
  const int P = a.P = (L + M - 1) / M;
  const int pitch = a.pitch = (L + 3) / 4 * 4;
  a.tab.resize((size_t)nu * 5 * pitch);
  a.GE.resize((size_t)nu * P * 2 * P);
  std::vector<char> okv(nu, 1);
  bt_parallel(nu, [&](int u) {
    okv[u] = bt_chunk_line(&a.lo[(size_t)u * L], &a.dg[(size_t)u * L], &a.hi[(size_t)u * L], L, M, pitch, false,
                           &a.tab[(size_t)u * 5 * pitch], &a.GE[(size_t)u * P * 2 * P]);
  });
  for (int u = 0; u < nu; ++u)
    if (!okv[u]) return false;
  a.band = bt_band(a.GE.data(), nu, P, 1e-16);
  if (want_il) {
    // chunk-interleaved copy [nu][5][M/2][P][2]: rows 2t, 2t+1 of chunk p at [.., t, p, :]; rows past the end of the
    // line: 1/piv = 1, everything else 0
    a.tab_il.assign((size_t)nu * 5 * P * M, 0.0);
    for (int u = 0; u < nu; ++u)
      for (int pl = 0; pl < 5; ++pl)
        for (int p = 0; p < P; ++p)
          for (int t = 0; t < M; ++t) {
            const int r = p * M + t;
            const double val = r < L ? a.tab[((size_t)u * 5 + pl) * pitch + r] : (pl == HS2_T_INV ? 1.0 : 0.0);
            a.tab_il[((((size_t)u * 5 + pl) * (M / 2) + t / 2) * P + p) * 2 + (t & 1)] = val;
          }
  }
  if (want_utab) {
    // the most common full-chunk table (weighted by the lines that use it) and where it applies
    const int n_full = L / M;
    a.utab.assign((size_t)5 * M, 0.0);
    a.ucode.assign((size_t)nu * P, 0);
    if (n_full > 0) {
      std::map<std::vector<double>, int64_t> votes;
      std::vector<double> key((size_t)5 * M);
      auto chunk_key = [&](int u, int p) {
        for (int pl = 0; pl < 5; ++pl)
          for (int t = 0; t < M; ++t) key[(size_t)pl * M + t] = a.tab[((size_t)u * 5 + pl) * pitch + p * M + t];
      };
      for (int u = 0; u < nu; ++u)
        for (int p = 0; p < n_full; ++p) {
          chunk_key(u, p);
          votes[key] += weight[u];
        }
      int64_t best = -1;
      for (auto &kv : votes)
        if (kv.second > best) best = kv.second, a.utab = kv.first;
      for (int u = 0; u < nu; ++u)
        for (int p = 0; p < n_full; ++p) {
          chunk_key(u, p);
          a.ucode[(size_t)u * P + p] = memcmp(key.data(), a.utab.data(), key.size() * sizeof(double)) == 0;
        }
    }
  }
  if (want_xw && L == P * M && L >= 3 * M) {
    // ghost-uniform lines (kernels_xw.cu): rows 1..L-2 identical (a, b, c), row 0 = (0, b + a, c), row L-1 =
    // (a, b + c, 0) to 4 ulp - the infinite constant-coefficient line with mirrored ghost neighbours
    a.xw_code.assign(nu, 0);
    std::vector<std::vector<double>> hdrs(nu), ges(nu);
    const double eps = 4 * 2.220446049250313e-16;
    bt_parallel(nu, [&](int u) {
      const double *lo = &a.lo[(size_t)u * L], *dg = &a.dg[(size_t)u * L], *hi = &a.hi[(size_t)u * L];
      const double aa = lo[1], bb = dg[1], cc = hi[1];
      bool ok = lo[0] == 0.0 && hi[L - 1] == 0.0 && std::fabs(dg[0] - (bb + aa)) <= eps * std::fabs(bb) &&
                std::fabs(dg[L - 1] - (bb + cc)) <= eps * std::fabs(bb);
      for (int r = 1; r < L && ok; ++r) ok = lo[r] == aa;
      for (int r = 0; r < L - 1 && ok; ++r) ok = hi[r] == cc;
      for (int r = 1; r < L - 1 && ok; ++r) ok = dg[r] == bb;
      if (!ok) return;
      std::vector<double> l2(L, aa), d2(L, bb), h2(L, cc), tab((size_t)5 * pitch);
      ges[u].resize((size_t)P * 2 * P);
      if (!bt_chunk_line(l2.data(), d2.data(), h2.data(), L, M, pitch, true, tab.data(), ges[u].data())) return;
      std::vector<double> &h = hdrs[u];
      h.assign(XW_HDR, 0.0);
      for (int pl = 0; pl < 5; ++pl)
        for (int t = 0; t < M; ++t) h[(size_t)pl * M + t] = tab[(size_t)pl * pitch + M + t];   // an interior chunk
      h[(size_t)HS2_T_F * M] = 0.0;
      double *bt = &h[(size_t)5 * M];
      for (int k = M - 2; k >= 0; --k) bt[k] = h[(size_t)HS2_T_S * M + k] - h[(size_t)HS2_T_CP * M + k] * bt[k + 1];
      h[(size_t)6 * M] = 1.0 / (1.0 + bt[0]);
      a.xw_code[u] = 1;
    });
    int band = 0;
    bool any = false;
    for (int u = 0; u < nu; ++u)
      if (a.xw_code[u]) {
        band = std::max(band, bt_band(ges[u].data(), 1, P, 1e-16));
        any = true;
      }
    a.xw_band = any ? band : 0;
    const int w = 2 * a.xw_band + 1;
    const size_t stride = XW_HDR + (size_t)w * P * 2;
    a.xw.assign((size_t)nu * stride, 0.0);
    for (int u = 0; u < nu; ++u) {
      if (!a.xw_code[u]) continue;
      double *dst = &a.xw[(size_t)u * stride];
      std::copy(hdrs[u].begin(), hdrs[u].end(), dst);
      for (int d = 0; d < w; ++d)
        for (int p = 0; p < P; ++p) {
          const int q = p + d - a.xw_band;
          if (q < 0 || q >= P) continue;
          dst[XW_HDR + ((size_t)d * P + p) * 2] = ges[u][(size_t)p * 2 * P + 2 * q];
          dst[XW_HDR + ((size_t)d * P + p) * 2 + 1] = ges[u][(size_t)p * 2 * P + 2 * q + 1];
        }
    }
  }
  return true;
}

static void bt_parallel(int n, const std::function<void(int)> &fn) {
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)std::min<unsigned>(hw ? hw : 1, 16);
  if (n < 4 * nt) nt = 1;
  if (nt <= 1) {
    for (int i = 0; i < n; ++i) fn(i);
    return;
  }
  std::vector<std::thread> th;
  for (int w = 0; w < nt; ++w)
    th.emplace_back([&, w]() {
      for (int i = w; i < n; i += nt) fn(i);
    });
  for (auto &t : th) t.join();
}

// Rows per chunk for a line of length L (same policy as heatsim2_b200/plan.py choose_chunk: M = 8 with up to 16
// chunks, 16 or 32 with up to 32; the largest M <= pref with at least 4 chunks, else the largest valid one;
// lines of 16..512 cells on the x axis, a multiple of 16: 16 (the TMA-fed x kernels).  0: whole-line fallback.
static int bt_choose_chunk(int L, int axis, int pref, bool x_tma, bool x_warp) {
  if (axis == 0 && x_tma && L % 16 == 0 && L >= 16 && L <= 512) return 16;
  if (axis == 0 && x_warp && L == 1024) return 16;     // two warps per line (kernels_xw.cu)
  const int Ms[3] = {8, 16, 32}, caps[3] = {16, 32, 32};
  int best_good = 0, smallest = 0, first_valid = 0;
  for (int q = 0; q < 3; ++q) {
    const int P = (L + Ms[q] - 1) / Ms[q];
    if (P > caps[q]) continue;
    if (!first_valid) first_valid = Ms[q];
    if (Ms[q] <= pref) {
      if (!smallest) smallest = Ms[q];
      if (P >= 4) best_good = Ms[q];
    }
  }
  return best_good ? best_good : (smallest ? smallest : first_valid);
}

struct hs2_owned {   // device buffers a built plan owns
  std::vector<void *> bufs;
  BtAxis axis[3];
  std::vector<uint32_t> h_line_id[3];
  ~hs2_owned() {
    for (void *b : bufs) cudaFree(b);
  }
};

template <typename T>
static int bt_upload(hs2_owned *own, const std::vector<T> &v, const T **out) {
  *out = nullptr;
  if (v.empty()) return HS2_OK;
  void *d = nullptr;
  HS2_CUDA_CHECK(cudaMalloc(&d, v.size() * sizeof(T)));
  own->bufs.push_back(d);
  HS2_CUDA_CHECK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = reinterpret_cast<const T *>(d);
  return HS2_OK;
}

// unique lines of one axis: line_id (host), representatives' rows -> BtAxis lo/dg/hi
template <typename CID>
static int bt_group_lines(const CID *d_cid, int axis, int L, int64_t n_lines, int64_t ny, int64_t nx,
                          const std::vector<double> &coef_raw, int n_classes, BtAxis &a, std::vector<uint32_t> &line_id,
                          std::vector<int64_t> &weight) {
  // row signature of every class on this axis: (lo, dg, hi) of (I - 1/2 M^-1 L_axis)
  const int gm = 1 + 2 * axis, gp = 2 + 2 * axis;
  std::map<std::array<double, 3>, int32_t> sig;
  std::vector<int32_t> sub(n_classes);
  std::vector<std::array<double, 3>> urows;
  for (int c = 0; c < n_classes; ++c) {
    const double cap = coef_raw[(size_t)c * 8];
    const std::array<double, 3> row = {-0.5 * coef_raw[(size_t)c * 8 + gm] / cap,
                                       1.0 + 0.5 * (coef_raw[(size_t)c * 8 + gm] + coef_raw[(size_t)c * 8 + gp]) / cap,
                                       -0.5 * coef_raw[(size_t)c * 8 + gp] / cap};
    auto it = sig.find(row);
    if (it == sig.end()) {
      it = sig.emplace(row, (int32_t)urows.size()).first;
      urows.push_back(row);
    }
    sub[c] = it->second;
  }
  // 64-bit polynomial weights
  std::vector<uint64_t> pw(L);
  {
    uint64_t acc = 1;
    for (int r = 0; r < L; ++r) {
      pw[r] = acc;
      acc = acc * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    }
  }
  int32_t *d_sub = nullptr;
  uint64_t *d_pw = nullptr, *d_hash = nullptr;
  int64_t *d_first = nullptr;
  int32_t *d_reps = nullptr;
  uint32_t *d_lid = nullptr;
  int *d_bad = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_sub), cudaFree(d_pw), cudaFree(d_hash), cudaFree(d_first), cudaFree(d_reps), cudaFree(d_lid), cudaFree(d_bad);
  };
#define BT_CHECK(call)                                                                                     \
  do {                                                                                                     \
    cudaError_t e__ = (call);                                                                              \
    if (e__ != cudaSuccess) {                                                                              \
      hs2_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);          \
      cleanup();                                                                                           \
      return HS2_E_CUDA;                                                                                   \
    }                                                                                                      \
  } while (0)
  BT_CHECK(cudaMalloc(&d_sub, sizeof(int32_t) * n_classes));
  BT_CHECK(cudaMalloc(&d_pw, sizeof(uint64_t) * L));
  BT_CHECK(cudaMalloc(&d_hash, sizeof(uint64_t) * n_lines));
  BT_CHECK(cudaMalloc(&d_bad, sizeof(int)));
  BT_CHECK(cudaMemcpy(d_sub, sub.data(), sizeof(int32_t) * n_classes, cudaMemcpyHostToDevice));
  BT_CHECK(cudaMemcpy(d_pw, pw.data(), sizeof(uint64_t) * L, cudaMemcpyHostToDevice));
  const int threads = 256;
  const int64_t work = axis == 0 ? n_lines * 32 : n_lines;
  const unsigned blocks = (unsigned)((work + threads - 1) / threads);
  bt_hash_kernel<CID><<<blocks, threads>>>(d_cid, d_sub, d_pw, axis, L, n_lines, ny, nx, d_hash);
  BT_CHECK(cudaGetLastError());
  std::vector<uint64_t> hash(n_lines);
  BT_CHECK(cudaMemcpy(hash.data(), d_hash, sizeof(uint64_t) * n_lines, cudaMemcpyDeviceToHost));
  // groups in order of first appearance
  std::unordered_map<uint64_t, uint32_t> ids;
  ids.reserve(1024);
  std::vector<int64_t> first;
  line_id.resize(n_lines);
  weight.clear();
  for (int64_t l = 0; l < n_lines; ++l) {
    auto it = ids.find(hash[l]);
    if (it == ids.end()) {
      it = ids.emplace(hash[l], (uint32_t)first.size()).first;
      first.push_back(l);
      weight.push_back(0);
    }
    line_id[l] = it->second;
    weight[it->second]++;
  }
  const int nu = (int)first.size();
  BT_CHECK(cudaMalloc(&d_first, sizeof(int64_t) * nu));
  BT_CHECK(cudaMalloc(&d_reps, sizeof(int32_t) * (size_t)nu * L));
  BT_CHECK(cudaMalloc(&d_lid, sizeof(uint32_t) * n_lines));
  BT_CHECK(cudaMemcpy(d_first, first.data(), sizeof(int64_t) * nu, cudaMemcpyHostToDevice));
  BT_CHECK(cudaMemcpy(d_lid, line_id.data(), sizeof(uint32_t) * n_lines, cudaMemcpyHostToDevice));
  BT_CHECK(cudaMemset(d_bad, 0, sizeof(int)));
  bt_gather_kernel<CID><<<(unsigned)(((int64_t)nu * L + threads - 1) / threads), threads>>>(d_cid, d_sub, d_first, axis, L, nu, ny,
                                                                                             nx, d_reps);
  BT_CHECK(cudaGetLastError());
  bt_verify_kernel<CID><<<blocks, threads>>>(d_cid, d_sub, d_lid, d_reps, axis, L, n_lines, ny, nx, d_bad);
  BT_CHECK(cudaGetLastError());
  int bad = 0;
  BT_CHECK(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<int32_t> reps((size_t)nu * L);
  BT_CHECK(cudaMemcpy(reps.data(), d_reps, sizeof(int32_t) * reps.size(), cudaMemcpyDeviceToHost));
  cleanup();
#undef BT_CHECK
  if (bad) {
    hs2_set_error("hs2_plan_build: 64-bit line-hash collision on axis %d (lines of one group differ)", axis);
    return HS2_E_INVALID;
  }
  a.L = L, a.nu = nu;
  a.lo.resize((size_t)nu * L), a.dg.resize((size_t)nu * L), a.hi.resize((size_t)nu * L);
  for (size_t e = 0; e < reps.size(); ++e) {
    const auto &row = urows[reps[e]];
    a.lo[e] = row[0], a.dg[e] = row[1], a.hi[e] = row[2];
  }
  return HS2_OK;
}

template <typename CID>
static int bt_build(const hs2_build_desc *b, hs2_owned *own, hs2_plan_desc *d) {
  const int64_t nz = b->nz, ny = b->ny, nx = b->nx;
  const CID *cid = reinterpret_cast<const CID *>(b->d_class_id);
  const bool slab = b->d_class_id_global != nullptr;
  std::vector<double> raw(b->h_class_coef, b->h_class_coef + (size_t)b->n_classes * 8);
  for (int c = 0; c < b->n_classes; ++c)
    HS2_REQUIRE(raw[(size_t)c * 8] != 0.0 && std::isfinite(raw[(size_t)c * 8]), "hs2_plan_build: class %d has capacity term %g", c,
                raw[(size_t)c * 8]);
  // closed outer faces (of the global grid for a slab)
  {
    const double *d_raw = nullptr;
    int rc = bt_upload(own, raw, &d_raw);
    if (rc) return rc;
    int *d_out = nullptr;
    HS2_CUDA_CHECK(cudaMalloc((void **)&d_out, 6 * sizeof(int)));
    own->bufs.push_back(d_out);
    HS2_CUDA_CHECK(cudaMemset(d_out, 0, 6 * sizeof(int)));
    const CID *gid = slab ? reinterpret_cast<const CID *>(b->d_class_id_global) : cid;
    const int64_t gz = slab ? b->nz_global : nz;
    const int64_t m = std::max(ny * nx, std::max(gz * nx, gz * ny));
    bt_closed_kernel<CID><<<(unsigned)((m + 255) / 256), 256>>>(gid, d_raw, gz, ny, nx, d_out);
    HS2_CUDA_CHECK(cudaGetLastError());
    int open[6];
    HS2_CUDA_CHECK(cudaMemcpy(open, d_out, sizeof(open), cudaMemcpyDeviceToHost));
    static const char *const names[6] = {"z-min", "z-max", "y-min", "y-max", "x-min", "x-max"};
    for (int f = 0; f < 6; ++f)
      HS2_REQUIRE(!open[f], "Equation exceeds bounds of domain on the %s face. Are external boundaries set correctly?", names[f]);
  }
  // scaled coefficient rows the kernels read
  std::vector<double> scaled((size_t)b->n_classes * HS2_COEF_STRIDE);
  for (int c = 0; c < b->n_classes; ++c) {
    const double cap = raw[(size_t)c * 8];
    for (int q = 0; q < 6; ++q) scaled[(size_t)c * 8 + q] = raw[(size_t)c * 8 + 1 + q] / cap;
    scaled[(size_t)c * 8 + 6] = raw[(size_t)c * 8 + 7] / cap;
    scaled[(size_t)c * 8 + 7] = cap;
  }
  int rc = bt_upload(own, scaled, &d->d_class_coef);
  if (rc) return rc;
  d->nz = nz, d->ny = ny, d->nx = nx;
  d->n_classes = b->n_classes;
  d->class_id_bytes = b->class_id_bytes;
  d->d_class_id = b->d_class_id;
  d->device = b->device;
  d->flags = b->flags;
  d->z_chunk0 = d->z_chunks_global = 0;
  const bool x_tma = !(b->flags & HS2_FLAG_X_FOLD);
  const bool x_warp = !(b->flags & (HS2_FLAG_X_FOLD | HS2_FLAG_X_PATCH)) && b->n_classes <= 64;
  for (int axis = 0; axis < 3; ++axis) {
    BtAxis &a = own->axis[axis];
    std::vector<int64_t> weight;
    const bool glob = slab && axis == 2;
    const int L = axis == 0 ? (int)nx : (axis == 1 ? (int)ny : (int)(glob ? b->nz_global : nz));
    const int64_t n_lines = axis == 0 ? nz * ny : (axis == 1 ? nz * nx : ny * nx);
    rc = bt_group_lines<CID>(glob ? reinterpret_cast<const CID *>(b->d_class_id_global) : cid, axis, L, n_lines, ny, nx, raw,
                             b->n_classes, a, own->h_line_id[axis], weight);
    if (rc) return rc;
    int M = b->chunk[axis];
    if (sizeof(CID) == 4) M = -1;      // per-cell classes: whole-line kernels, Thomas factors only
    if (M == 0) M = glob ? 0 : bt_choose_chunk(L, axis, 32, x_tma, x_warp);
    if (M < 0) M = 0;
    HS2_REQUIRE(M == 0 || M == 8 || M == 16 || M == 32, "hs2_plan_build: chunk[%d] = %d (0, 8, 16 or 32)", axis, M);
    if (glob) HS2_REQUIRE(M > 0 && nz % M == 0, "hs2_plan_build: slab thickness %lld is not a multiple of the z chunk %d", (long long)nz, M);
    a.M = M;
    const bool utab = !glob && M > 0 && (b->utab_axes & (1 << axis));
    const bool xw = axis == 0 && M == 16 && (nx == 1024 || nx == 512 || nx == 256) && x_warp;
    HS2_REQUIRE(bt_axis_tables(a, weight, !glob, axis == 0 && M > 0, utab, xw), "hs2_plan_build: singular interface system on axis %d",
                axis);
    hs2_axis_tables &t = d->axis[axis];
    memset(&t, 0, sizeof(t));
    if ((rc = bt_upload(own, own->h_line_id[axis], &t.d_line_id))) return rc;
    if ((rc = bt_upload(own, a.lu, &t.d_lu))) return rc;
    if ((rc = bt_upload(own, a.tab, &t.d_tab))) return rc;
    if ((rc = bt_upload(own, a.GE, &t.d_GE))) return rc;
    if ((rc = bt_upload(own, a.tab_il, &t.d_tab_il))) return rc;
    if ((rc = bt_upload(own, a.ucode, &t.d_ucode))) return rc;
    if ((rc = bt_upload(own, a.xw, &t.d_xw_tab))) return rc;
    if ((rc = bt_upload(own, a.xw_code, &t.d_xw_code))) return rc;
    t.h_utab = a.utab.empty() ? nullptr : a.utab.data();
    t.n_unique = a.nu, t.chunk = M, t.n_chunks = a.P, t.pitch = a.pitch, t.band = a.band, t.xw_band = a.xw_band;
    if (glob) {
      d->z_chunk0 = (int32_t)(b->k0 / M);
      d->z_chunks_global = a.P;
    }
  }
  return HS2_OK;
}

void hs2_owned_free(hs2_owned *o) { delete o; }

extern "C" {

int hs2_plan_build(const hs2_build_desc *b, hs2_plan **out) {
  HS2_REQUIRE(b && out, "hs2_plan_build: NULL argument");
  *out = nullptr;
  HS2_REQUIRE(b->nz > 0 && b->ny > 0 && b->nx > 0, "hs2_plan_build: empty grid");
  HS2_REQUIRE(b->class_id_bytes == 1 || b->class_id_bytes == 2 || b->class_id_bytes == 4, "hs2_plan_build: class_id_bytes must be 1, 2 or 4");
  HS2_REQUIRE(b->n_classes > 0 && (b->class_id_bytes == 4 || b->n_classes <= (b->class_id_bytes == 1 ? 256 : 65536)),
              "hs2_plan_build: n_classes %d out of range", b->n_classes);
  HS2_REQUIRE(b->class_id_bytes != 4 || !b->d_class_id_global, "hs2_plan_build: 4-byte class ids are not supported in z-slab plans");
  HS2_REQUIRE(b->d_class_id && b->h_class_coef, "hs2_plan_build: NULL class tables");
  HS2_REQUIRE(b->nx < ((int64_t)1 << 30) && b->ny < ((int64_t)1 << 30) && b->nz < ((int64_t)1 << 30), "hs2_plan_build: grid too large");
  if (b->d_class_id_global)
    HS2_REQUIRE(b->k0 >= 0 && b->k0 + b->nz <= b->nz_global, "hs2_plan_build: slab [%lld, %lld) outside the global grid of %lld planes",
                (long long)b->k0, (long long)(b->k0 + b->nz), (long long)b->nz_global);
  int prev = 0;
  HS2_CUDA_CHECK(cudaGetDevice(&prev));
  HS2_CUDA_CHECK(cudaSetDevice(b->device));
  hs2_owned *own = new (std::nothrow) hs2_owned;
  if (!own) {
    hs2_set_error("hs2_plan_build: out of host memory");
    return HS2_E_NOMEM;
  }
  hs2_plan_desc d;
  memset(&d, 0, sizeof(d));
  int rc = b->class_id_bytes == 1 ? bt_build<uint8_t>(b, own, &d)
                                  : (b->class_id_bytes == 2 ? bt_build<uint16_t>(b, own, &d) : bt_build<uint32_t>(b, own, &d));
  if (!rc) rc = hs2_plan_create(&d, out);
  cudaSetDevice(prev);
  if (rc) {
    delete own;
    return rc;
  }
  (*out)->owned = own;
  return HS2_OK;
}

int hs2_plan_axis_info(const hs2_plan *plan, int axis, hs2_axis_info *info) {
  HS2_REQUIRE(plan && info && axis >= 0 && axis < 3, "hs2_plan_axis_info: bad argument");
  const hs2_axis_tables &t = plan->d.axis[axis];
  info->n_unique = t.n_unique, info->chunk = t.chunk, info->n_chunks = t.n_chunks, info->pitch = t.pitch;
  info->band = t.band, info->xw_band = t.xw_band;
  info->line_length = axis == 0 ? plan->d.nx : (axis == 1 ? plan->d.ny : (plan->d.z_chunks_global ? (int64_t)t.chunk * t.n_chunks : plan->d.nz));
  info->n_lines = axis == 0 ? plan->d.nz * plan->d.ny : (axis == 1 ? plan->d.nz * plan->d.nx : plan->d.ny * plan->d.nx);
  return HS2_OK;
}

int64_t hs2_plan_copy_table(const hs2_plan *plan, int axis, int which, void *h_dst, int64_t capacity) {
  if (!plan || axis < 0 || axis > 2 || !plan->owned) {
    hs2_set_error("hs2_plan_copy_table: plan was not made by hs2_plan_build");
    return HS2_E_INVALID;
  }
  const BtAxis &a = plan->owned->axis[axis];
  const void *src = nullptr;
  int64_t bytes = 0;
  switch (which) {
    case HS2_TAB_LINE_ID: src = plan->owned->h_line_id[axis].data(), bytes = plan->owned->h_line_id[axis].size() * 4; break;
    case HS2_TAB_ROWS_LO: src = a.lo.data(), bytes = a.lo.size() * 8; break;
    case HS2_TAB_ROWS_DG: src = a.dg.data(), bytes = a.dg.size() * 8; break;
    case HS2_TAB_ROWS_HI: src = a.hi.data(), bytes = a.hi.size() * 8; break;
    case HS2_TAB_LU: src = a.lu.data(), bytes = a.lu.size() * 8; break;
    case HS2_TAB_CHUNK: src = a.tab.data(), bytes = a.tab.size() * 8; break;
    case HS2_TAB_GE: src = a.GE.data(), bytes = a.GE.size() * 8; break;
    case HS2_TAB_CHUNK_IL: src = a.tab_il.data(), bytes = a.tab_il.size() * 8; break;
    case HS2_TAB_UTAB: src = a.utab.data(), bytes = a.utab.size() * 8; break;
    case HS2_TAB_UCODE: src = a.ucode.data(), bytes = a.ucode.size(); break;
    case HS2_TAB_XW: src = a.xw.data(), bytes = a.xw.size() * 8; break;
    case HS2_TAB_XW_CODE: src = a.xw_code.data(), bytes = a.xw_code.size(); break;
    default: hs2_set_error("hs2_plan_copy_table: unknown table %d", which); return HS2_E_INVALID;
  }
  if (h_dst && bytes > 0) {
    if (capacity < bytes) {
      hs2_set_error("hs2_plan_copy_table: buffer of %lld bytes, table has %lld", (long long)capacity, (long long)bytes);
      return HS2_E_INVALID;
    }
    memcpy(h_dst, src, (size_t)bytes);
  }
  return bytes;
}

// The table algebra alone, on host arrays (no device needed): lines [nu][L] -> chunk tables, as hs2_plan_build
// computes them.  tab [nu][5][pitch], GE [nu][P][2P] (pitch = L rounded up to 4, P = ceil(L / M)); returns the band.
int hs2_tables_chunk(const double *lo, const double *dg, const double *hi, int nu, int L, int M, int ghost, double *tab, double *GE) {
  HS2_REQUIRE(lo && dg && hi && tab && GE && nu > 0 && L > 0 && (M == 8 || M == 16 || M == 32), "hs2_tables_chunk: bad argument");
  const int P = (L + M - 1) / M, pitch = (L + 3) / 4 * 4;
  for (int u = 0; u < nu; ++u)
    HS2_REQUIRE(bt_chunk_line(lo + (size_t)u * L, dg + (size_t)u * L, hi + (size_t)u * L, L, M, pitch, ghost != 0,
                              tab + (size_t)u * 5 * pitch, GE + (size_t)u * P * 2 * P),
                "hs2_tables_chunk: singular interface system (line %d)", u);
  return bt_band(GE, nu, P, 1e-16);
}

}  // extern "C"

// kernels_v1.cu - straightforward global-memory implementation of the three
// ADI stages.  Correct for every grid size; used as the fall-back when a line
// does not fit the shared-memory tile kernels (kernels_tile.cu) and as an
// independent cross-check of them in the tests.
//
// Stage algebra (delta form of Douglas-Gunn; the reference's
// three stage equations are heatsim2/alternatingdirection_c_pyx.pyx:482-493):
//   x: (I - 1/2 M^-1 Lx) d1 = M^-1 [(Lx+Ly+Lz) T + D s]
//   y: (I - 1/2 M^-1 Ly) d2 = d1
//   z: (I - 1/2 M^-1 Lz) d3 = d2 ,  T_new = T + d3
// Rows are scaled by 1/M so stages y and z need no per-cell coefficient at all:
// their tridiagonal factors come from the per-unique-line tables.
#include "hs2_common.cuh"

namespace {

struct SrcTable {
  int n;
  uint8_t idx[8];
  double val[8];
};

template <typename CID>
__global__ void __launch_bounds__(256)
rhs_kernel(const double *__restrict__ T, double *__restrict__ W,
           const CID *__restrict__ cid, const double *__restrict__ coef,
           const uint8_t *__restrict__ vol, SrcTable st,
           const double *__restrict__ dense, const double *__restrict__ halo_lo,
           const double *__restrict__ halo_hi, int64_t nz, int64_t ny, int64_t nx) {
  const int64_t n = nz * ny * nx;
  const int64_t plane = ny * nx;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx % nx;
    const int64_t j = (idx / nx) % ny;
    const int64_t k = idx / plane;
    const double *c = coef + (int64_t)cid[idx] * HS2_COEF_STRIDE;
    const double tc = T[idx];
    const double xm = i > 0 ? T[idx - 1] : tc;
    const double xp = i < nx - 1 ? T[idx + 1] : tc;
    const double ym = j > 0 ? T[idx - nx] : tc;
    const double yp = j < ny - 1 ? T[idx + nx] : tc;
    const double zm = k > 0 ? T[idx - plane] : (halo_lo ? halo_lo[idx] : tc);
    const double zp = k < nz - 1 ? T[idx + plane]
                                 : (halo_hi ? halo_hi[idx - (nz - 1) * plane] : tc);
    double r = c[0] * (xm - tc) + c[1] * (xp - tc);
    r += c[2] * (ym - tc) + c[3] * (yp - tc);
    r += c[4] * (zm - tc) + c[5] * (zp - tc);
    double s = dense ? dense[idx] : 0.0;
    if (vol && st.n) {
      const uint8_t v = vol[idx];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < st.n && st.idx[q] == v) s += st.val[q];
    }
    r += c[6] * s;
    W[idx] = r;
  }
}

// One thread per line, in-place Thomas with precomputed factors.
// base(line) = (line / inner) * outer + (line % inner); element r at base + r*stride.
__global__ void __launch_bounds__(128)
thomas_kernel(double *__restrict__ W, const double *__restrict__ Tin,
              double *__restrict__ Tout, const uint32_t *__restrict__ line_id,
              const double *__restrict__ lu, int64_t n_lines, int64_t L,
              int64_t stride, int64_t inner, int64_t outer) {
  const int64_t line = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (line >= n_lines) return;
  const int64_t base = (line / inner) * outer + (line % inner);
  const double4 *f = reinterpret_cast<const double4 *>(lu) + (int64_t)line_id[line] * L;
  double *w = W + base;
  double prev = 0.0;
  for (int64_t r = 0; r < L; ++r) {
    const double4 c = f[r];  // {inv, lo*inv, hi*inv, 0}
    const double u = fma(-c.y, prev, w[r * stride] * c.x);
    w[r * stride] = u;
    prev = u;
  }
  double next = 0.0;
  for (int64_t r = L - 1; r >= 0; --r) {
    const double4 c = f[r];
    const double x = fma(-c.z, next, w[r * stride]);
    next = x;
    if (Tout)
      Tout[base + r * stride] = Tin[base + r * stride] + x;
    else
      w[r * stride] = x;
  }
}

int make_src_table(const hs2_source *src, SrcTable *st) {
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

}  // namespace

int hs2_v1_sweep_x(hs2_plan *p, const double *T, double *W, const hs2_source *src,
                   const double *halo_lo, const double *halo_hi, cudaStream_t st) {
  SrcTable tab;
  int rc = make_src_table(src, &tab);
  if (rc) return rc;
  p->last_kernel[0] = HS2_K_WHOLE_LINE;
  const hs2_plan_desc &d = p->d;
  const int threads = 256;
  int64_t blocks64 = (p->n + threads - 1) / threads;
  const int64_t cap = (int64_t)p->sm_count * 32;
  const int blocks = (int)(blocks64 < cap ? blocks64 : cap);
  const uint8_t *vol = (src && tab.n) ? src->d_vol_elements : nullptr;
  const double *dense = src ? src->d_dense : nullptr;
  if (d.class_id_bytes == 1)
    rhs_kernel<uint8_t><<<blocks, threads, 0, st>>>(T, W, (const uint8_t *)d.d_class_id, d.d_class_coef, vol, tab,
                                                    dense, halo_lo, halo_hi, d.nz, d.ny, d.nx);
  else if (d.class_id_bytes == 4)
    rhs_kernel<uint32_t><<<blocks, threads, 0, st>>>(T, W, (const uint32_t *)d.d_class_id, d.d_class_coef, vol, tab,
                                                     dense, halo_lo, halo_hi, d.nz, d.ny, d.nx);
  else
    rhs_kernel<uint16_t><<<blocks, threads, 0, st>>>(T, W, (const uint16_t *)d.d_class_id, d.d_class_coef, vol, tab,
                                                     dense, halo_lo, halo_hi, d.nz, d.ny, d.nx);
  HS2_CUDA_CHECK(cudaGetLastError());
  const int64_t n_lines = d.nz * d.ny;
  thomas_kernel<<<(unsigned)((n_lines + 127) / 128), 128, 0, st>>>(W, nullptr, nullptr, d.axis[0].d_line_id, d.axis[0].d_lu,
                                                                   n_lines, d.nx, 1, 1, d.nx);
  HS2_CUDA_CHECK(cudaGetLastError());
  return HS2_OK;
}

int hs2_v1_sweep_y(hs2_plan *p, double *W, cudaStream_t st) {
  p->last_kernel[1] = HS2_K_WHOLE_LINE;
  const hs2_plan_desc &d = p->d;
  const int64_t n_lines = d.nz * d.nx;
  thomas_kernel<<<(unsigned)((n_lines + 127) / 128), 128, 0, st>>>(W, nullptr, nullptr, d.axis[1].d_line_id, d.axis[1].d_lu,
                                                                   n_lines, d.ny, d.nx, d.nx, d.ny * d.nx);
  HS2_CUDA_CHECK(cudaGetLastError());
  return HS2_OK;
}

int hs2_v1_sweep_z(hs2_plan *p, const double *T, double *Tout, double *W, cudaStream_t st) {
  p->last_kernel[2] = HS2_K_WHOLE_LINE;
  const hs2_plan_desc &d = p->d;
  const int64_t n_lines = d.ny * d.nx;
  thomas_kernel<<<(unsigned)((n_lines + 127) / 128), 128, 0, st>>>(W, T, Tout, d.axis[2].d_line_id, d.axis[2].d_lu, n_lines,
                                                                   d.nz, d.ny * d.nx, n_lines, 0);
  HS2_CUDA_CHECK(cudaGetLastError());
  return HS2_OK;
}

// cabi.cu - extern "C" surface declared in include/hs2_b200.h.
#include <stdarg.h>
#include <new>

#include "hs2_common.cuh"

static thread_local char g_err[512] = "";

void hs2_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static bool has_source(const hs2_source *src) {
  return src && ((src->h_value && src->d_vol_elements) || src->d_dense);
}

extern "C" {

int hs2_abi_version(void) { return HS2_ABI_VERSION; }

const char *hs2_last_error(void) { return g_err; }

int hs2_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(hs2_axis_tables);
    case 1: return (int)sizeof(hs2_plan_desc);
    case 2: return (int)sizeof(hs2_source);
    case 3: return (int)sizeof(hs2_build_desc);
    case 4: return (int)sizeof(hs2_axis_info);
  }
  return -1;
}

int hs2_plan_create(const hs2_plan_desc *desc, hs2_plan **out) {
  HS2_REQUIRE(desc && out, "hs2_plan_create: NULL argument");
  *out = nullptr;
  HS2_REQUIRE(desc->nz > 0 && desc->ny > 0 && desc->nx > 0, "hs2_plan_create: empty grid %lld x %lld x %lld",
              (long long)desc->nz, (long long)desc->ny, (long long)desc->nx);
  HS2_REQUIRE(desc->class_id_bytes == 1 || desc->class_id_bytes == 2 || desc->class_id_bytes == 4,
              "hs2_plan_create: class_id_bytes must be 1, 2 or 4");
  HS2_REQUIRE(desc->n_classes > 0 && (desc->class_id_bytes == 4 || desc->n_classes <= (desc->class_id_bytes == 1 ? 256 : 65536)),
              "hs2_plan_create: n_classes %d out of range for %d-byte ids", desc->n_classes, desc->class_id_bytes);
  HS2_REQUIRE(desc->class_id_bytes != 4 || desc->z_chunks_global == 0, "hs2_plan_create: 4-byte class ids are not supported in z-slab plans");
  HS2_REQUIRE(desc->d_class_id && desc->d_class_coef, "hs2_plan_create: NULL class tables");
  for (int a = 0; a < 3; ++a)
    HS2_REQUIRE(desc->axis[a].d_line_id && desc->axis[a].d_lu && desc->axis[a].n_unique > 0,
                "hs2_plan_create: missing line table for axis %d", a);
  int ndev = 0;
  HS2_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  HS2_REQUIRE(desc->device >= 0 && desc->device < ndev, "hs2_plan_create: device %d not present (%d visible)",
              desc->device, ndev);
  cudaDeviceProp prop;
  HS2_CUDA_CHECK(cudaGetDeviceProperties(&prop, desc->device));
  HS2_REQUIRE(prop.major == 10, "hs2_plan_create: built for sm_100a, device %d is sm_%d%d", desc->device, prop.major,
              prop.minor);
  hs2_plan *p = new (std::nothrow) hs2_plan;
  if (!p) {
    hs2_set_error("hs2_plan_create: out of host memory");
    return HS2_E_NOMEM;
  }
  p->d = *desc;
  // 4-byte class ids (per-cell equations, e.g. a curvature map: up to 2^31 classes): the coefficient rows do not fit
  // shared memory and every line is its own class - the whole-line global-memory kernels run all three sweeps
  if (desc->class_id_bytes == 4) p->d.flags |= HS2_FLAG_FORCE_FALLBACK;
  p->owned = nullptr;
  p->graph_cache = nullptr;
  p->n = desc->nz * desc->ny * desc->nx;
  p->sm_count = prop.multiProcessorCount;
  p->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  p->last_kernel[0] = p->last_kernel[1] = p->last_kernel[2] = HS2_K_NONE;
  for (int a = 0; a < 3; ++a) {
    const hs2_axis_tables &ax = desc->axis[a];
    memset(&p->utab[a], 0, sizeof(UTab));
    p->has_utab[a] = ax.h_utab && ax.d_ucode && ax.chunk > 0 && ax.chunk <= 32;
    if (p->has_utab[a])
      for (int pl = 0; pl < HS2_T_PLANES; ++pl)
        for (int t = 0; t < ax.chunk; ++t) p->utab[a].v[pl][t] = ax.h_utab[pl * ax.chunk + t];
    p->d.axis[a].h_utab = nullptr;   // the caller's buffer is not referenced after this call
  }
  *out = p;
  return HS2_OK;
}

int hs2_plan_destroy(hs2_plan *plan) {
  if (plan) hs2_graph_cache_free(plan);
  if (plan && plan->owned) hs2_owned_free(plan->owned);
  delete plan;
  return HS2_OK;
}

int hs2_plan_launches_per_step(const hs2_plan *plan) {
  if (!plan) return 0;
  return ((hs2_tile_xt_supported(plan) || hs2_tile_xf_supported(plan)) ? 1 : 2) + 2;
}

int hs2_plan_x_kernel(const hs2_plan *plan) {
  if (!plan) return -1;
  if (hs2_tile_xw_supported(plan)) return HS2_XK_WARP;
  if (hs2_tile_xt_supported(plan)) return HS2_XK_TMA;
  if (!hs2_tile_xf_supported(plan)) return HS2_XK_WHOLE_LINE;
  return HS2_XK_FOLD;
}

int hs2_plan_last_kernel(const hs2_plan *plan, int axis) {
  if (!plan || axis < 0 || axis > 2) return -1;
  return plan->last_kernel[axis];
}

const char *hs2_kernel_name(int code) {
  static const char *const names[] = {"none", "whole-line", "tile", "tile-tma", "tile-tma-512", "tile-cpasync",
                                      "tile-cpasync-512", "x-fold", "z-slab", "x-tma", "x-warp", "tile-rows", "tile-rows-512"};
  return (code >= 0 && code < (int)(sizeof(names) / sizeof(names[0]))) ? names[code] : "?";
}

int hs2_sweep_x(hs2_plan *plan, const double *d_T_in, double *d_work, const hs2_source *src, const double *d_halo_lo,
                const double *d_halo_hi, void *stream) {
  HS2_REQUIRE(plan && d_T_in && d_work, "hs2_sweep_x: NULL argument");
  HS2_REQUIRE(d_T_in != d_work, "hs2_sweep_x: d_work must not alias d_T_in");
  if (!has_source(src) && hs2_tile_xw_supported(plan)) {
    bool done = false;
    int rc = hs2_tile_sweep_xw(plan, d_T_in, d_work, d_halo_lo, d_halo_hi, 0, (cudaStream_t)stream, &done);
    if (rc || done) return rc;
  }
  if (hs2_tile_xt_supported(plan)) {
    bool done = false;
    int rc = hs2_tile_sweep_xt(plan, d_T_in, d_work, src, d_halo_lo, d_halo_hi, 0, (cudaStream_t)stream, &done);
    if (rc || done) return rc;
  }
  if (hs2_tile_xf_supported(plan))
    return hs2_tile_sweep_xf(plan, d_T_in, d_work, src, d_halo_lo, d_halo_hi, 0, (cudaStream_t)stream);
  return hs2_v1_sweep_x(plan, d_T_in, d_work, src, d_halo_lo, d_halo_hi, (cudaStream_t)stream);
}

int hs2_sweep_x_part(hs2_plan *plan, const double *d_T_in, double *d_work, const hs2_source *src,
                     const double *d_halo_lo, const double *d_halo_hi, int part, void *stream) {
  HS2_REQUIRE(plan && d_T_in && d_work, "hs2_sweep_x_part: NULL argument");
  HS2_REQUIRE(d_T_in != d_work, "hs2_sweep_x_part: d_work must not alias d_T_in");
  HS2_REQUIRE(part == HS2_X_INTERIOR || part == HS2_X_BOUNDARY, "hs2_sweep_x_part: part must be HS2_X_INTERIOR or HS2_X_BOUNDARY");
  if (!has_source(src) && hs2_tile_xw_supported(plan) && plan->d.nz >= 3) {
    bool done = false;
    int rc = hs2_tile_sweep_xw(plan, d_T_in, d_work, d_halo_lo, d_halo_hi, part, (cudaStream_t)stream, &done);
    if (rc || done) return rc;
  }
  if (hs2_tile_xt_supported(plan) && plan->d.nz >= 3) {
    bool done = false;
    int rc = hs2_tile_sweep_xt(plan, d_T_in, d_work, src, d_halo_lo, d_halo_hi, part, (cudaStream_t)stream, &done);
    if (rc || done) return rc;
  }
  if (!hs2_tile_xf_supported(plan) || plan->d.nz < 3) {
    // kernels without plane ranges: everything happens in the boundary call
    if (part == HS2_X_INTERIOR) return HS2_OK;
    return hs2_sweep_x(plan, d_T_in, d_work, src, d_halo_lo, d_halo_hi, stream);
  }
  return hs2_tile_sweep_xf(plan, d_T_in, d_work, src, d_halo_lo, d_halo_hi, part, (cudaStream_t)stream);
}

int hs2_sweep_y(hs2_plan *plan, double *d_work, void *stream) {
  HS2_REQUIRE(plan && d_work, "hs2_sweep_y: NULL argument");
  if (hs2_tile_supported(plan, 1)) return hs2_tile_sweep_y(plan, d_work, (cudaStream_t)stream);
  return hs2_v1_sweep_y(plan, d_work, (cudaStream_t)stream);
}

int hs2_sweep_z(hs2_plan *plan, const double *d_T_in, double *d_T_out, double *d_work, void *stream) {
  HS2_REQUIRE(plan && d_T_in && d_T_out && d_work, "hs2_sweep_z: NULL argument");
  HS2_REQUIRE(plan->d.z_chunks_global == 0, "hs2_sweep_z: slab plans use hs2_sweep_z_forward/backward");
  if (hs2_tile_supported(plan, 2)) return hs2_tile_sweep_z(plan, d_T_in, d_T_out, d_work, (cudaStream_t)stream);
  return hs2_v1_sweep_z(plan, d_T_in, d_T_out, d_work, (cudaStream_t)stream);
}

int hs2_sweep_z_forward(hs2_plan *plan, double *d_work, double *d_Y, int64_t line0, int64_t n_lines, void *stream) {
  HS2_REQUIRE(plan && d_work && d_Y, "hs2_sweep_z_forward: NULL argument");
  return hs2_zdist(plan, 0, d_work, nullptr, nullptr, d_Y, line0, n_lines, 0, nullptr, false, (cudaStream_t)stream);
}

int hs2_sweep_z_forward_push(hs2_plan *plan, double *d_work, double *d_Y, int n_peers, const uint64_t *peer_Y,
                             void *stream) {
  HS2_REQUIRE(plan && d_work && d_Y, "hs2_sweep_z_forward_push: NULL argument");
  return hs2_zdist(plan, 0, d_work, nullptr, nullptr, d_Y, 0, plan->d.ny * plan->d.nx, n_peers, peer_Y, true,
                   (cudaStream_t)stream);
}

int hs2_sweep_z_forward_push_cols(hs2_plan *plan, double *d_work, double *d_Y, int64_t line0, int64_t n_lines, int n_peers,
                                  const uint64_t *peer_Y, void *stream) {
  HS2_REQUIRE(plan && d_work && d_Y, "hs2_sweep_z_forward_push_cols: NULL argument");
  return hs2_zdist(plan, 0, d_work, nullptr, nullptr, d_Y, line0, n_lines, n_peers, peer_Y, true, (cudaStream_t)stream);
}

int hs2_sweep_z_backward_cols(hs2_plan *plan, const double *d_T_in, double *d_T_out, double *d_work, const double *d_Yall,
                              int64_t line0, int64_t n_lines, void *stream) {
  HS2_REQUIRE(plan && d_T_in && d_T_out && d_work && d_Yall, "hs2_sweep_z_backward_cols: NULL argument");
  return hs2_zdist(plan, 1, d_work, d_T_in, d_T_out, const_cast<double *>(d_Yall), line0, n_lines, 0, nullptr, true,
                   (cudaStream_t)stream);
}

int hs2_sweep_z_backward(hs2_plan *plan, const double *d_T_in, double *d_T_out, double *d_work, const double *d_Yall,
                         int64_t line0, int64_t n_lines, void *stream) {
  HS2_REQUIRE(plan && d_T_in && d_T_out && d_work && d_Yall, "hs2_sweep_z_backward: NULL argument");
  return hs2_zdist(plan, 1, d_work, d_T_in, d_T_out, const_cast<double *>(d_Yall), line0, n_lines, 0, nullptr, false,
                   (cudaStream_t)stream);
}

int hs2_sweep_z_fused(hs2_plan *plan, const double *d_T_in, double *d_T_out, double *d_work, double *d_Yall, int n_peers,
                      const uint64_t *peer_Y, double timeout_s, int *d_status, void *stream) {
  HS2_REQUIRE(plan && d_T_in && d_T_out && d_work && d_Yall, "hs2_sweep_z_fused: NULL argument");
  return hs2_zfused(plan, d_work, d_T_in, d_T_out, d_Yall, n_peers, peer_Y, timeout_s, d_status, nullptr, (cudaStream_t)stream);
}

int hs2_sweep_z_fused_tile_lines(hs2_plan *plan) {
  if (!plan) return HS2_E_INVALID;
  int lines = 0;
  int rc = hs2_zfused(plan, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0.0, nullptr, &lines, nullptr);
  return rc ? rc : lines;
}

int hs2_peer_fill_empty(void *d_ptr, int64_t n_doubles, void *stream) {
  HS2_REQUIRE(d_ptr && n_doubles >= 0, "hs2_peer_fill_empty: bad argument");
  return hs2_fill_empty(d_ptr, n_doubles, (cudaStream_t)stream);
}

int hs2_step(hs2_plan *plan, const double *d_T_in, double *d_T_out, double *d_work, const hs2_source *src,
             const double *d_halo_lo, const double *d_halo_hi, void *stream) {
  int rc = hs2_sweep_x(plan, d_T_in, d_work, src, d_halo_lo, d_halo_hi, stream);
  if (rc) return rc;
  rc = hs2_sweep_y(plan, d_work, stream);
  if (rc) return rc;
  HS2_REQUIRE(plan->d.z_chunks_global == 0, "hs2_step: slab plans are stepped with hs2_sweep_x/y + hs2_sweep_z_forward/backward");
  return hs2_sweep_z(plan, d_T_in, d_T_out, d_work, stream);
}

}  // extern "C"

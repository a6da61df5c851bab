/* hs2_b200.h - C ABI of the B200-native ADI Crank-Nicolson time step.
 *
 * This is the drop-in boundary for heatsim2's hot path.  Every entry point
 * names the reference interface it replaces (paths relative to the
 * isuthermography/heatsim2 tree).  Plain pointers and sizes only; all device
 * buffers are owned by the caller (the Python host layer hands in
 * torch.Tensor.data_ptr() values), the plan owns nothing but a validated copy
 * of its descriptor.  Kernels are launched on the CUDA stream passed in
 * (a cudaStream_t cast to void*; NULL = legacy default stream) and never
 * synchronise the device.
 *
 * Error convention: every function returns 0 on success or a negative
 * HS2_E_* code; hs2_last_error() gives the message (thread-local).  Nothing
 * calls abort()/exit() - the reference's assert/exit(1) paths
 * (heatsim2/alternatingdirection_c.c:1-3,160-163) become error returns.
 */
#ifndef HS2_B200_H
#define HS2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HS2_ABI_VERSION 6

#define HS2_OK 0
#define HS2_E_INVALID (-1) /* bad argument / unsupported shape               */
#define HS2_E_CUDA (-2)    /* a CUDA runtime call failed                      */
#define HS2_E_NOMEM (-3)

/* Per equation-class coefficient row (8 doubles), already divided by the
 * capacity term M = rho*c/dt of the class (M = 1, g = 0, D = 0 for a
 * TEMPERATURE_FIXED class):
 *   [0] gx-/M  [1] gx+/M  [2] gy-/M  [3] gy+/M  [4] gz-/M  [5] gz+/M
 *   [6] D/M    [7] M
 * g are the face conductances (W/m^3/K) that heatsim2/crank_nicolson.pyx:
 * 359-363 forms symbolically per cell; D is the volumetric-source weight
 * (heatsim2/alternatingdirection_c.c:119-121).                              */
#define HS2_COEF_STRIDE 8

/* Per unique tridiagonal line, per row: {1/pivot, lower/pivot, upper/pivot, 0}
 * of the Thomas factorisation of (I - 1/2 M^-1 L_axis) - the information
 * heatsim2/tridiag.pyx:9-43 (tridiaglu) stores as Lmat/Umat, 72 B per row
 * there, 32 B per row of a *unique* line here.                              */
#define HS2_LU_STRIDE 4

/* Partitioned (SPIKE-type) solve: a line of length L is cut into n_chunks
 * chunks of `chunk` rows (last one may be shorter); each chunk is eliminated
 * independently and the chunk interfaces are recovered from a precomputed
 * dense operator.  Tables per unique line (heatsim2_b200/plan.py,
 * chunk_factors, documents the algebra):
 *   d_tab [n_unique][HS2_T_PLANES][pitch]  planes 1/piv, lo/piv, c, s, hi/piv
 *                                          (chunk-local factorisation)
 *   d_GE  [n_unique][n_chunks][2*n_chunks] row of the inverse interface
 *                                          operator giving the chunk's last x
 *   d_tab_il [n_unique][HS2_T_PLANES][chunk/2][n_chunks][2]  (x axis only, may
 *                                          be NULL) the same planes with the
 *                                          chunks interleaved: rows 2t, 2t+1 of
 *                                          chunk p at ((plane*chunk/2 + t)*
 *                                          n_chunks + p)*2; rows past the end
 *                                          of the line hold 1/piv = 1, rest 0
 *   h_utab [HS2_T_PLANES][chunk]           HOST copy of the most common chunk
 *                                          table of this axis (may be NULL): the
 *                                          interior chunks of lines with constant
 *                                          coefficients are bit-identical; the
 *                                          kernels receive it by value (constant
 *                                          bank) and read it without loads
 *   d_ucode [n_unique][n_chunks] (u8)      1 where chunk p of unique line u
 *                                          equals h_utab bit for bit, else 0
 *   d_xw_tab [n_unique][128 + (2*xw_band+1)*64]  (x axis only, may be NULL; the warp-per-line
 *                                          kernel, lines of 32 chunks of 16) per unique line
 *                                          with constant coefficients and closed ends: [0..79] the
 *                                          common chunk table (planes as d_tab, 16 rows each),
 *                                          [80..95] b (correction of the first chunk), [96] 1/(1+b_0),
 *                                          then [2*xw_band+1][32 chunks][2] rows of the inverse
 *                                          interface operator with mirrored ghost neighbours,
 *                                          diagonal-relative (entry d of chunk p multiplies chunk
 *                                          p + d - xw_band; 0 outside the line)
 *   d_xw_code [n_unique] (u8)              1 where d_xw_tab describes the line, else 0
 * chunk == 0 means the tables are absent and the whole-line path is used.   */
#define HS2_T_INV 0
#define HS2_T_F 1
#define HS2_T_C 2
#define HS2_T_S 3
#define HS2_T_CP 4
#define HS2_T_PLANES 5

typedef struct hs2_axis_tables {
  const uint32_t *d_line_id; /* device [n_lines]: unique-line id of each line  */
  const double *d_lu;        /* device [n_unique][L][HS2_LU_STRIDE]            */
  const double *d_tab;
  const double *d_GE;
  const double *d_tab_il;
  const double *h_utab;      /* HOST pointer, read at launch time, or NULL       */
  const uint8_t *d_ucode;    /* device, or NULL                                  */
  int32_t n_unique;
  int32_t chunk;
  int32_t n_chunks;
  int32_t pitch;             /* doubles per table plane (even, >= L)           */
  int32_t band;              /* half-width (in chunks) of d_GE rows worth applying */
  int32_t xw_band;           /* the same for the rows in d_xw_tab                  */
  const double *d_xw_tab;    /* device, or NULL                                  */
  const uint8_t *d_xw_code;  /* device, or NULL                                  */
} hs2_axis_tables;

/* Axis numbering used by every per-axis array: 0 = x (contiguous), 1 = y,
 * 2 = z (slowest).  Line numbering: x-lines k*ny+j, y-lines k*nx+i,
 * z-lines j*nx+i.                                                            */
typedef struct hs2_plan_desc {
  int64_t nz, ny, nx;          /* local grid (a z-slab in multi-GPU runs)     */
  int32_t n_classes;
  int32_t class_id_bytes;      /* 1 (u8), 2 (u16) or 4 (u32: whole-line kernels only) */
  const void *d_class_id;      /* device, [nz][ny][nx]                        */
  const double *d_class_coef;  /* device, [n_classes][HS2_COEF_STRIDE]        */
  hs2_axis_tables axis[3];
  int32_t device;              /* CUDA device ordinal the buffers live on     */
  int32_t flags;               /* HS2_FLAG_*                                  */
  /* z-slab decomposition (one slab per GPU).  When z_chunks_global > 0 the
   * plan's nz planes are chunks [z_chunk0, z_chunk0 + nz/chunk) of global
   * z-lines made of z_chunks_global chunks; axis[2] then describes the GLOBAL
   * line: d_tab planes are indexed by global row (pitch >= global nz), d_GE is
   * [n_unique][z_chunks_global][2*z_chunks_global], and nz must be a multiple
   * of axis[2].chunk.  0 / 0 for a single-GPU plan.                           */
  int32_t z_chunk0;
  int32_t z_chunks_global;
} hs2_plan_desc;

#define HS2_FLAG_FORCE_FALLBACK 1 /* use the whole-line global-memory kernels */
#define HS2_FLAG_NO_UTAB 4        /* ignore h_utab / d_ucode (every chunk reads its factor tables) */
#define HS2_FLAG_X_FOLD 16        /* x sweep: keep the folded LSU-fed kernel even where the TMA-fed ones apply */
#define HS2_FLAG_X_PATCH 32       /* x sweep: keep the TMA-fed patch kernel where the warp-per-line kernel applies */

typedef struct hs2_plan hs2_plan;

/* Volumetric source of one step (replaces the per-call volumetric_array of
 * heatsim2/alternatingdirection_c_pyx.pyx:294-386).  Either field may be
 * NULL.  Table form: src(cell) = value[vol_elements(cell)] with a 256-entry
 * host table (IMPULSE / STEPPED / POINT_JOULES with scalar volume); dense
 * form: a full [nz][ny][nx] device array in W/m^3 (z-decaying impulse,
 * per-cell volumes).                                                          */
typedef struct hs2_source {
  const uint8_t *d_vol_elements; /* device [nz][ny][nx] or NULL               */
  const double *h_value;         /* host [256] or NULL                        */
  const double *d_dense;         /* device [nz][ny][nx] or NULL               */
} hs2_source;

int hs2_abi_version(void);
const char *hs2_last_error(void);
/* sizeof of the ABI structs as compiled (binding self-check): which = 0 hs2_axis_tables,
 * 1 hs2_plan_desc, 2 hs2_source, 3 hs2_build_desc, 4 hs2_axis_info; -1 for anything else */
int hs2_sizeof(int which);

/* replaces create_adi_step x3 + finalize (alternatingdirection_c.h:52,
 * alternatingdirection_c_pyx.pyx:212-282)                                     */
int hs2_plan_create(const hs2_plan_desc *desc, hs2_plan **out);
/* The reference's C boundary builds a stage by itself: create_adi_step + one add_equation per cell
 * (heatsim2/alternatingdirection_c.h:50-54, alternatingdirection_c.c:56-199), then tridiaglu
 * (heatsim2/tridiag.pyx:9-43).  hs2_plan_build is that step here: from the per-cell class ids and the per-class
 * coefficient rows alone it derives every table of hs2_plan_desc (unique lines per axis, Thomas factors,
 * partitioned-solve tables, interface operators, the x kernels' extra tables) in native host and device code;
 * the plan owns what it allocates.  No host-language table code is needed to drive the library.
 *   h_class_coef  HOST [n_classes][8], as parsed from the equations, un-scaled:
 *                 M = rho*c/dt, gx-, gx+, gy-, gy+, gz-, gz+ (face conductances), D (source weight)
 *   chunk[a]      rows per chunk of the partitioned solve on axis a: 0 = choose, -1 = whole-line kernels only,
 *                 8 / 16 / 32 = forced
 *   utab_axes     bit a set: build the common-chunk table of axis a (h_utab / d_ucode)
 *   d_class_id_global, nz_global, k0: z-slab of a larger grid (multi-GPU): the z tables describe the global lines
 * Errors: HS2_E_INVALID with "Equation exceeds bounds of domain ..." where a conductance points out of the grid
 * (the reference exits there, alternatingdirection_c.c:160-163).                                             */
typedef struct hs2_build_desc {
  int64_t nz, ny, nx;
  int32_t n_classes;
  int32_t class_id_bytes;        /* 1 (u8), 2 (u16) or 4 (u32: per-cell equations, whole-line kernels) */
  const void *d_class_id;        /* device [nz][ny][nx]; must outlive the plan  */
  const double *h_class_coef;    /* host [n_classes][8]                         */
  int32_t device;
  int32_t flags;                 /* HS2_FLAG_*                                  */
  int32_t chunk[3];
  int32_t utab_axes;
  const void *d_class_id_global; /* device [nz_global][ny][nx] or NULL          */
  int64_t nz_global, k0;
} hs2_build_desc;
int hs2_plan_build(const hs2_build_desc *desc, hs2_plan **out);

/* what hs2_plan_build (or the caller of hs2_plan_create) chose for one axis    */
typedef struct hs2_axis_info {
  int64_t line_length, n_lines;
  int32_t n_unique, chunk, n_chunks, pitch, band, xw_band;
} hs2_axis_info;
int hs2_plan_axis_info(const hs2_plan *plan, int axis, hs2_axis_info *info);

/* Host copy of a table of a plan made by hs2_plan_build (inspection, tests, and the multi-GPU driver's set-up):
 * returns the table's size in bytes (h_dst may be NULL to query it) or a negative error code.              */
#define HS2_TAB_LINE_ID 0   /* u32 [n_lines]                        */
#define HS2_TAB_ROWS_LO 1   /* f64 [n_unique][L] sub-diagonal       */
#define HS2_TAB_ROWS_DG 2   /*                   diagonal           */
#define HS2_TAB_ROWS_HI 3   /*                   super-diagonal     */
#define HS2_TAB_LU 4        /* d_lu                                 */
#define HS2_TAB_CHUNK 5     /* d_tab                                */
#define HS2_TAB_GE 6        /* d_GE                                 */
#define HS2_TAB_CHUNK_IL 7  /* d_tab_il                             */
#define HS2_TAB_UTAB 8      /* h_utab                               */
#define HS2_TAB_UCODE 9     /* d_ucode                              */
#define HS2_TAB_XW 10       /* d_xw_tab                             */
#define HS2_TAB_XW_CODE 11  /* d_xw_code                            */
int64_t hs2_plan_copy_table(const hs2_plan *plan, int axis, int which, void *h_dst, int64_t capacity);

/* The table algebra alone on host arrays (no device): unique lines lo/dg/hi [nu][L] -> tab [nu][5][pitch]
 * (pitch = L rounded up to 4) and GE [nu][P][2P], P = ceil(L / M); ghost != 0: mirrored neighbours at both ends
 * (d_xw_tab).  Returns the interface band (>= 0) or a negative error code.                                   */
int hs2_tables_chunk(const double *lo, const double *dg, const double *hi, int nu, int L, int M, int ghost,
                     double *tab, double *GE);

/* replaces delete_adi_step (alternatingdirection_c.h:50)                      */
int hs2_plan_destroy(hs2_plan *plan);
/* number of kernels one hs2_step launches with this plan                     */
int hs2_plan_launches_per_step(const hs2_plan *plan);

/* which kernel a whole-grid hs2_sweep_x of this plan runs: whole-line
 * global-memory fallback (rhs + Thomas), the folded tile kernel, the
 * TMA-fed patch kernel or the TMA-fed warp-per-line kernel (source-free steps;
 * steps with a volumetric source run the patch kernel)                       */
#define HS2_XK_WHOLE_LINE 0
#define HS2_XK_FOLD 1
#define HS2_XK_TMA 2
#define HS2_XK_WARP 3
int hs2_plan_x_kernel(const hs2_plan *plan);

/* Which kernel variant the LAST sweep along `axis` (0 = x, 1 = y, 2 = z) of
 * this plan launched: one of HS2_K_* (0 before the first sweep).  Lets a
 * caller - and the parity tests - verify that a given grid ran on the code
 * path it was meant to exercise.  hs2_kernel_name gives the code's name.      */
#define HS2_K_NONE 0
#define HS2_K_WHOLE_LINE 1     /* kernels_v1.cu: one thread per line, global memory      */
#define HS2_K_TILE 2           /* strided_sweep: register tile, one block per tile       */
#define HS2_K_TILE_TMA 3       /* strided_sweep_tma: persistent, next tile by TMA        */
#define HS2_K_TILE_TMA_BIG 4   /* the same with 512-thread blocks (17..32 chunks)        */
#define HS2_K_TILE_CPASYNC 5   /* persistent, next tile by 16-byte cp.async              */
#define HS2_K_TILE_CPASYNC_BIG 6
#define HS2_K_X_FOLD 7         /* sweep_xf_kernel                                        */
#define HS2_K_Z_SLAB 8         /* z_forward / z_backward of a slab plan                  */
#define HS2_K_X_TMA 9          /* sweep_xt_kernel: patches staged by TMA, chunk-layout RHS */
#define HS2_K_X_WARP 10        /* sweep_xw_kernel: one warp per line, interfaces by shuffle */
#define HS2_K_TILE_ROWS 11     /* persistent, next tile by one bulk copy per row         */
#define HS2_K_TILE_ROWS_BIG 12
int hs2_plan_last_kernel(const hs2_plan *plan, int axis);
const char *hs2_kernel_name(int code);

/* One ADI time step = run_adi_steps (alternatingdirection_c_pyx.pyx:287-416).
 *   d_T_in   [nz][ny][nx]  field at t - dt/2
 *   d_T_out  [nz][ny][nx]  field at t + dt/2 (may alias d_T_in)
 *   d_work   [nz][ny][nx]  scratch (stage increments)
 *   src      source of this step or NULL
 *   d_halo_lo / d_halo_hi: [ny][nx] planes of T_in just below k=0 / above
 *            k=nz-1 owned by the neighbouring slab, or NULL at a domain face. */
int hs2_step(hs2_plan *plan, const double *d_T_in, double *d_T_out,
             double *d_work, const hs2_source *src, const double *d_halo_lo,
             const double *d_halo_hi, void *stream);

/* Many steps per call, field resident in HBM (the reference's demos call run_adi_steps once per step from Python and
 * copy the whole field back to index it: demos/steelonfoam.py:132-143, demos/ktest.py).  Runs the source-free steps
 * first_step .. first_step + nsteps - 1: step n reads d_T_a for even n and d_T_b for odd n and writes the other
 * buffer.  After every step with (n + 1) % every == 0 the observation of hs2_observe is appended to the record
 * arrays at row *d_counter (a 64-bit device word the caller owns and sets to the first free row; it is incremented
 * on the device after each record, so that the launches can be replayed):
 *   d_probe_rec   [rows][n_probes]  or NULL      d_surface_rec [rows][ny][nx]  or NULL
 * use_graph != 0: a unit of steps is captured once into a CUDA graph (kept by the plan for the next call with the
 * same buffers) and replayed - no per-launch cost on small grids; 0: plain launches.  Asynchronous on `stream`.  */
int hs2_run_steps(hs2_plan *plan, double *d_T_a, double *d_T_b, double *d_work, int64_t first_step, int64_t nsteps,
                  int every, const int64_t *d_probe_cells, int n_probes, double *d_probe_rec, double *d_surface_rec,
                  double dz, uint64_t *d_counter, int use_graph, void *stream);
/* Observation of one field on the device, one launch:
 *   d_probe_out[q]      = T[d_probe_cells[q]]   (flat cell index (k*ny + j)*nx + i), q < n_probes
 *   d_surface_out[j][i] = the insulated z-min surface temperature estimated from planes 0 and 1,
 *                         heatsim2/surface_temperature.py:4-36 (same operations in the same order)
 * Either output may be NULL.                                                                                       */
int hs2_observe(hs2_plan *plan, const double *d_T, const int64_t *d_probe_cells, int n_probes, double *d_probe_out,
                double *d_surface_out, double dz, void *stream);

/* The three stages separately (multi-GPU drivers interleave communication).
 * hs2_sweep_x: d_work = (M - Lx/2)^-1 [(Lx+Ly+Lz) T_in + D src]
 * hs2_sweep_y: d_work = (M - Ly/2)^-1 M d_work
 * hs2_sweep_z: d_T_out = d_T_in + (M - Lz/2)^-1 M d_work   (local z lines)   */
int hs2_sweep_x(hs2_plan *plan, const double *d_T_in, double *d_work,
                const hs2_source *src, const double *d_halo_lo,
                const double *d_halo_hi, void *stream);
/* hs2_sweep_x in two launches, so that a slab's halo exchange can travel while
 * the planes that do not need it are being solved: HS2_X_INTERIOR = planes
 * 1..nz-2 (the halo pointers are not read), HS2_X_BOUNDARY = planes 0 and nz-1.
 * Both calls together equal one hs2_sweep_x.                                  */
#define HS2_X_INTERIOR 1
#define HS2_X_BOUNDARY 2
int hs2_sweep_x_part(hs2_plan *plan, const double *d_T_in, double *d_work,
                     const hs2_source *src, const double *d_halo_lo,
                     const double *d_halo_hi, int part, void *stream);
int hs2_sweep_y(hs2_plan *plan, double *d_work, void *stream);
int hs2_sweep_z(hs2_plan *plan, const double *d_T_in, double *d_T_out,
                double *d_work, void *stream);

/* z-sweep of a slab plan, in two halves around one exchange (there is no
 * reference counterpart: heatsim2 is single-process).  Both calls work on the
 * z-lines [line0, line0 + n_lines) (line = j*nx + i), so that a caller can
 * pipeline: while one range of lines is being exchanged the next one is solved.
 *   forward : eliminate the local chunks of d_work (d_work is left untouched)
 *             and write their (y_first, y_last) pairs to d_Y [2*nz/chunk][n_lines]
 *   <caller delivers d_Y of every slab that lies within the interface band,
 *    in slab order, into d_Yall [2*z_chunks_global][n_lines] - NCCL over NVLink>
 *   backward: repeat the elimination, interface values from d_Yall,
 *             back-substitution, d_T_out = d_T_in + increment               */
int hs2_sweep_z_forward(hs2_plan *plan, double *d_work, double *d_Y,
                        int64_t line0, int64_t n_lines, void *stream);
int hs2_sweep_z_backward(hs2_plan *plan, const double *d_T_in, double *d_T_out,
                         double *d_work, const double *d_Yall,
                         int64_t line0, int64_t n_lines, void *stream);

/* NVLink peer-memory variant of the exchange (one process per GPU, all on one
 * node).  hs2_peer_alloc returns zero-initialised device memory and a 64-byte
 * CUDA IPC handle; another process maps it with hs2_peer_open.  The z-interface
 * values are then written by the producing kernel itself:
 * hs2_sweep_z_forward_push = hs2_sweep_z_forward over all lines that also
 * stores this slab's 2*nz/chunk rows at peer_Y[i] (device addresses inside the
 * peers' mapped buffers, row pitch ny*nx doubles), i < n_peers <=
 * HS2_MAX_Z_PEERS.  Ordering between GPUs is by 64-bit flags in peer memory:
 * hs2_flag_signal stores `value` (release, system scope) to every listed flag
 * after all earlier work of the stream; hs2_flag_wait holds the stream until
 * every listed (local) flag is >= value, or sets *d_status = 1 after
 * timeout_s seconds instead of hanging.  Flag lists are host arrays of device
 * addresses, at most 16 entries.  hs2_copy_async = cudaMemcpyAsync between
 * any two mapped device addresses (halo planes over NVLink).                */
#define HS2_MAX_Z_PEERS 6
int hs2_peer_alloc(int64_t bytes, void **d_ptr, void *handle64);
int hs2_peer_open(const void *handle64, void **d_ptr);
int hs2_peer_close(void *d_ptr);
int hs2_peer_free(void *d_ptr);
int hs2_sweep_z_forward_push(hs2_plan *plan, double *d_work, double *d_Y,
                             int n_peers, const uint64_t *peer_Y, void *stream);
/* The same two halves on a RANGE of z-lines [line0, line0 + n_lines) with the interface rows addressed by the
 * line's number in the slab (row pitch ny*nx doubles, as in the mailboxes): the peer-memory step pipelines
 * forward(range 1) behind the wait for range 0's rows, and backward(range 0) behind the wait for range 1's. */
int hs2_sweep_z_forward_push_cols(hs2_plan *plan, double *d_work, double *d_Y, int64_t line0, int64_t n_lines,
                                  int n_peers, const uint64_t *peer_Y, void *stream);
int hs2_sweep_z_backward_cols(hs2_plan *plan, const double *d_T_in, double *d_T_out, double *d_work,
                              const double *d_Yall, int64_t line0, int64_t n_lines, void *stream);
/* Both halves AND the exchange in one persistent kernel: a block eliminates a tile of lines, stores its interface
 * rows into d_Yall (own rows) and the peers' mailboxes (peer_Y as for hs2_sweep_z_forward_push), eliminates the next
 * tile while the rows travel, then back-substitutes the first one from registers: the increment is read once
 * (24 B per cell instead of 32) and no wait sits between two launches.  The interface rows are their own flags:
 * every slot a peer writes holds an "empty" NaN pattern (hs2_peer_fill_empty, once, after hs2_peer_alloc) until the
 * value arrives, the kernel spins (bounded, *d_status = 1 on expiry) on the slots it needs and empties them again
 * after use.  d_Yall as for hs2_sweep_z_backward_cols; hs2_sweep_z_fused_tile_lines: lines per tile (information). */
int hs2_sweep_z_fused(hs2_plan *plan, const double *d_T_in, double *d_T_out, double *d_work, double *d_Yall,
                      int n_peers, const uint64_t *peer_Y, double timeout_s, int *d_status, void *stream);
int hs2_sweep_z_fused_tile_lines(hs2_plan *plan);
int hs2_peer_fill_empty(void *d_ptr, int64_t n_doubles, void *stream);
int hs2_flag_signal(const uint64_t *flag_ptrs, int n, uint64_t value, void *stream);
int hs2_flag_wait(const uint64_t *flag_ptrs, int n, uint64_t value,
                  double timeout_s, int *d_status, void *stream);
int hs2_copy_async(void *d_dst, const void *d_src, int64_t bytes, void *stream);

/* Drop-ins for heatsim2/tridiag.pyx on device arrays.
 * hs2_tridiag_lu    = tridiaglu   (:9-43):  A[n][3] -> L[n][3], U[n][3]
 * hs2_tridiag_solve = tridiagsolve(:46-69): x = U^-1 L^-1 b, one chain of n
 * rows, evaluated as a parallel scan of the two first-order recurrences.
 * d_scratch: at least hs2_tridiag_scratch_bytes(n) bytes.                    */
int64_t hs2_tridiag_scratch_bytes(int64_t n);
int hs2_tridiag_lu(int64_t n, const double *d_A, double *d_L, double *d_U,
                   void *d_scratch, void *stream);
int hs2_tridiag_solve(int64_t n, const double *d_L, const double *d_U,
                      const double *d_b, double *d_x, void *d_scratch,
                      void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HS2_B200_H */

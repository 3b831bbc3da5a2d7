/*
 * treemorph_nn — C ABI of the B200-native nearest-cylinder label + offset path.
 *
 * This is the drop-in boundary for ONE hot path of RobinDanek/Extracting-Tree-Morphology-From-Point-Clouds:
 * the point -> QSM-cylinder search.  The reference has no FFI for it (it is a chain of ~70 ATen
 * calls per 1024-point batch); its "operator API" is three Python callables per module, and each
 * entry point below names the reference lines it replaces (paths relative to the reference root):
 *
 *   closest_cylinder_cuda_batch         PreProcessing/LabelGenerationCuda.py:20-111   (variant A)
 *                                       Modules/Projection.py:19-115                  (variant B)
 *   generate_offset_cloud_cuda_batched  PreProcessing/LabelGenerationCuda.py:113-135
 *                                       Modules/Projection.py:117-144
 *
 * and, either side of that path:
 *
 *   cylinder_proximity_based_segmentation  Modules/Pipeline/QSMFittingDepthFirst.py:1006-1094  (tm_proximity_flags_host)
 *   compute_*_ckdtree / add_features       Modules/Features.py:111-229                         (tm_knn_covariance, tm_radius_count)
 *   noiseGeneration                        PreProcessing/NoiseDataGeneration.py:14-106         (tm_noise_cloud)
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs; no torch / C++ types cross this boundary;
 *   - every pointer is a DEVICE pointer on the handle's device unless the name ends in `_host`;
 *   - the caller owns inputs and outputs; the library owns only scratch memory inside the handle;
 *   - every call returns an int status (TM_OK == 0); tm_last_error() gives the message;
 *     no exceptions cross the boundary;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls are
 *     asynchronous with respect to the host unless stated otherwise;
 *   - one handle per (device, host thread); handles are not thread-safe;
 *   - there is NO CPU fallback: without a CUDA device tm_create() fails.
 *
 * The Python side binds this with ctypes and passes tensor.data_ptr() values
 * (extracting-tree-morphology-from-point-clouds_b200/binding.py); INTEGRATION.md shows the stub.
 */
#ifndef TREEMORPH_NN_H
#define TREEMORPH_NN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TM_ABI_VERSION 7

/* status codes */
#define TM_OK               0
#define TM_ERR_INVALID      1   /* bad argument (null pointer, negative size, unknown mode ...)      */
#define TM_ERR_NO_CYLINDERS 2   /* N > 0 but M == 0: the reference raises in argmin on an empty dim  */
#define TM_ERR_CUDA         3   /* a CUDA runtime call or kernel failed; see tm_last_error()          */
#define TM_ERR_NOMEM        4   /* device or pinned-host allocation failed                            */
#define TM_ERR_STATE        5   /* call order violated (e.g. tm_label_points before tm_set_cylinders) */

/* tm_params.mode */
#define TM_MODE_AUTO  0         /* grid when N*M is large enough to pay for it, else brute force      */
#define TM_MODE_BRUTE 1         /* exhaustive tiled kernel: every (point, cylinder) pair              */
#define TM_MODE_GRID  2         /* voxel-binned points + per-voxel cylinder tiles (exact pruning)     */

/* cloud element types for the host/record entry points */
#define TM_F32 0
#define TM_F64 1

typedef struct tm_handle tm_handle;

/*
 * The knobs that distinguish the reference's two copies of the kernel.
 *   variant A (LabelGenerationCuda.py): perp_atol 1e-6 (:51), norm_eps 0   (:58-60, unguarded)
 *   variant B (Projection.py):          perp_atol 1e-3 (:50), norm_eps 1e-8 (:60-63)
 */
typedef struct tm_params {
    float   perp_atol;       /* isclose(dot, 0, atol) tolerance                                       */
    float   norm_eps;        /* lower guard on ||rejection|| before the divide; 0 = unguarded         */
    int32_t move_to_mantle;  /* 1 = offset ends on the mantle / cap rim (reference behaviour);
                                0 = offset ends on the distance foot point (Projection.py:19 flag;
                                the reference's own False branch is shape-broken, :110)               */
    int32_t norm_fma;        /* which of ATen's two CPU roundings of norm() to mirror:
                                0 = strided xyz axis (cylinder tensors from DataFrame.values, the
                                    label_clouds / project_clouds path): sqrt((x*x + y*y) + z*z)
                                1 = contiguous xyz axis (C-ordered tensors, QSMFittingDepthFirst.py:1039,
                                    or M == 1): sqrt(fma(z,z, fma(y,y, x*x)))                         */
    int32_t mode;            /* TM_MODE_*                                                             */
    float   cell_size;       /* voxel edge in metres for TM_MODE_GRID; 0 = automatic                  */
    int32_t reserved[2];     /* must be 0                                                             */
} tm_params;

/* Counters of the most recent tm_label_points / tm_label_cloud_host call (tm_get_stats syncs). */
typedef struct tm_stats {
    uint64_t pairs_evaluated;   /* full (point, cylinder) evaluations in reference arithmetic, all kernels   */
    uint64_t cull_tests;        /* capsule lower-bound tests that decided whether a pair is evaluated        */
    uint64_t points_grid;       /* points certified by the near part of their own voxel's tile               */
    uint64_t points_far;        /* points certified by the far part of their own voxel's tile                */
    uint64_t points_ring;       /* points certified by the ball query over neighbouring voxels (few stragglers) */
    uint64_t points_tree;       /* points answered by the bounding-volume-hierarchy search (clutter, outside the grid) */
    uint64_t points_brute;      /* points answered by the exhaustive kernel (non-finite points, brute mode)  */
    uint64_t index_entries;     /* (voxel, cylinder) entries of the static voxel index                       */
    uint32_t voxels_occupied;   /* voxels holding at least one point                                         */
    uint32_t work_items;        /* occupied voxels (one tile descriptor each) processed by the tile kernel   */
    uint32_t mode_used;         /* TM_MODE_BRUTE or TM_MODE_GRID                                             */
    float    cell_size;         /* voxel edge actually used                                                  */
    float    reach;             /* D_max: a tile lists every cylinder within D_max of its voxel              */
    float    near_reach;        /* D_near: radius covered by the near part of a tile (<= reach)              */
    uint32_t grid_dim[3];       /* voxel grid extent                                                         */
    uint32_t launches;          /* kernels the call launched                                                 */
    uint64_t bound_tests;       /* closed-form distance estimates (point, tile entry) of the tile kernel     */
    uint64_t points_slow;       /* points the tile kernel settled with reference-order evaluations           */
    uint32_t lane_ops_per_bound;/* FP32 lane-operations of one estimate (the F of its roofline share)        */
    uint32_t reserved;
} tm_stats;

int tm_version(void);                                   /* returns TM_ABI_VERSION                     */
int tm_create(int device, tm_handle **out);             /* fails with TM_ERR_CUDA if no such device   */
int tm_destroy(tm_handle *h);
const char *tm_last_error(const tm_handle *h);          /* valid until the next call on h             */
const char *tm_status_string(int status);

/*
 * Cylinder preparation — replaces LabelGenerationCuda.py:121-123 / Projection.py:126-132:
 *   axis = end - start;  axis_length = ||axis||;  axis_unit = axis / max(axis_length, axis_eps)
 * start/end: (M,3) with element strides (row_stride, col_stride) — DataFrame.values tensors are
 * Fortran-ordered, i.e. strides (1, M).  out_axis_length (M,), out_axis_unit (M,3) row-major.
 * axis_eps: 0 (variant A) or 1e-8 (variant B).  norm_fma as in tm_params.
 */
int tm_prepare_cylinders(tm_handle *h,
                         const float *start, int64_t start_row_stride, int64_t start_col_stride,
                         const float *end, int64_t end_row_stride, int64_t end_col_stride,
                         int64_t m, float axis_eps, int32_t norm_fma,
                         float *out_axis_length, float *out_axis_unit, void *stream);

/*
 * Install the cylinder table (the cylinder arguments of closest_cylinder_cuda_batch,
 * LabelGenerationCuda.py:20 / Projection.py:19).  The values are used exactly as given
 * (axis_unit / axis_length are NOT recomputed).  Packs two float4 records per cylinder and
 * computes capsule AABBs; the static voxel index (per-voxel cylinder tiles) is built on the first
 * grid-mode call for a given cell size.
 * ids may be NULL (then id == row index).  Synchronises `stream` (it sizes the grid on the host).
 */
int tm_set_cylinders(tm_handle *h,
                     const float *start, int64_t start_row_stride, int64_t start_col_stride,
                     const float *axis_unit, int64_t unit_row_stride, int64_t unit_col_stride,
                     const float *axis_length, int64_t length_stride,
                     const float *radius, int64_t radius_stride,
                     const int32_t *ids, int64_t ids_stride,
                     int64_t m, void *stream);

/*
 * Multi-GPU: the search shards by POINTS, every rank needs the whole cylinder table and nothing else, so the
 * only collective is one broadcast of the table from the rank that read the QSM (the tensors built at
 * LabelGenerationCuda.py:117-123 / Projection.py:121-132); afterwards each rank labels its own rows with tm_label_points /
 * tm_label_cloud_host and no data-path collective follows.  NCCL is bound at run time (dlopen of libnccl.so.2).
 *
 *   one process per GPU:   rank 0 calls tm_comm_unique_id and ships the 128 bytes to the others (MPI, a file, torch's store);
 *                          every rank calls tm_comm_init_rank on its own handle, then tm_broadcast_cylinders (the cylinder
 *                          pointers / m are read on `root` only; strides in elements as for tm_set_cylinders).  On return the
 *                          table is installed on every rank, exactly as tm_set_cylinders would have installed it.
 *   one process, N GPUs:   tm_comm_init_all over N handles (one per device), tm_broadcast_cylinders_all with the source
 *                          pointers on handles[root_index]'s device.
 * Without a communicator (or with one rank) tm_broadcast_cylinders is tm_set_cylinders.  Both synchronise.
 */
#define TM_COMM_ID_BYTES 128
int tm_cylinder_count(const tm_handle *h, int64_t *m);           /* rows of the installed table */
int tm_comm_unique_id(void *id_out);
int tm_comm_init_rank(tm_handle *h, const void *id, int32_t nranks, int32_t rank);
int tm_comm_init_all(tm_handle **handles, int32_t ndev);
int tm_comm_destroy(tm_handle *h);
int tm_comm_info(const tm_handle *h, int32_t *rank, int32_t *nranks);
int tm_broadcast_cylinders(tm_handle *h,
                           const float *start, int64_t start_row_stride, int64_t start_col_stride,
                           const float *axis_unit, int64_t unit_row_stride, int64_t unit_col_stride,
                           const float *axis_length, int64_t length_stride,
                           const float *radius, int64_t radius_stride,
                           const int32_t *ids, int64_t ids_stride,
                           int64_t m, int32_t root, void *stream);
int tm_broadcast_cylinders_all(tm_handle **handles, int32_t ndev,
                               const float *start, int64_t start_row_stride, int64_t start_col_stride,
                               const float *axis_unit, int64_t unit_row_stride, int64_t unit_col_stride,
                               const float *axis_length, int64_t length_stride,
                               const float *radius, int64_t radius_stride,
                               const int32_t *ids, int64_t ids_stride,
                               int64_t m, int32_t root_index);

/*
 * closest_cylinder_cuda_batch for N device-resident points (LabelGenerationCuda.py:20-111,
 * Projection.py:19-115).  pts: fp32, row i at pts + i*row_stride (elements), xyz contiguous.
 * Outputs (any may be NULL): out_index (N,) winning ROW, out_id (N,) = ids[row] (:109),
 * out_dist (N,) (:88), out_offset (N,3) row-major (:106), out_radius (N,) radius of the winner.
 * N == 0 is a no-op.  M == 0 with N > 0 returns TM_ERR_NO_CYLINDERS.
 */
int tm_label_points(tm_handle *h, const float *pts, int64_t n, int64_t row_stride,
                    const tm_params *params,
                    int32_t *out_index, int32_t *out_id, float *out_dist, float *out_offset,
                    float *out_radius, void *stream);

/*
 * The (N,7) float64 record [x, y, z, ox, oy, oz, ID] of generate_offset_cloud_cuda_batched
 * (LabelGenerationCuda.py:114,131-133): xyz copied from the caller's cloud at ITS precision
 * (cloud_dtype TM_F32 / TM_F64, row stride in elements), offsets and ids widened to float64.
 */
int tm_assemble_records(tm_handle *h, const void *cloud, int32_t cloud_dtype, int64_t n,
                        int64_t cloud_row_stride, const float *offset, const int32_t *id,
                        double *out_records, void *stream);

/*
 * generate_offset_cloud_cuda_batched end to end for a HOST cloud (LabelGenerationCuda.py:113-135,
 * Projection.py:117-144): chunks the cloud, overlaps H2D copy / labelling / record assembly / D2H
 * on internal streams, and writes the (N,7) float64 record to out_records_host.
 * cloud_host: (N, >=3) TM_F32 or TM_F64, row stride in elements.  out_dist_host (N,) may be NULL.
 * Host buffers may be pageable; a pinned cloud avoids a staging copy.  Synchronous.
 */
int tm_label_cloud_host(tm_handle *h, const void *cloud_host, int32_t cloud_dtype, int64_t n,
                        int64_t cloud_row_stride, const tm_params *params,
                        double *out_records_host, float *out_dist_host);

/*
 * The same call writing the drivers' final record directly: rows of `row_doubles` (7..15) float64, the first seven as above,
 * the others filled with tail_values[0 .. row_doubles-7) — label_clouds / project_clouds without features append four
 * columns of ones to get the (N,11) layout TreeSet expects (LabelGenerationCuda.py:199-200, Projection.py:428-432).
 * out_rows_host may be a memory-mapped .npy file: the rows are written once, with streaming stores, where np.save would
 * otherwise copy them a second time (LabelGenerationCuda.py:203-205).  row_doubles == 7 is tm_label_cloud_host.
 */
int tm_label_cloud_host_wide(tm_handle *h, const void *cloud_host, int32_t cloud_dtype, int64_t n,
                             int64_t cloud_row_stride, const tm_params *params,
                             double *out_rows_host, int32_t row_doubles, const double *tail_values,
                             float *out_dist_host);

/*
 * Small-table fast path — the call pattern of cylinder_proximity_based_segmentation
 * (Modules/Pipeline/QSMFittingDepthFirst.py:1006-1094): thousands of calls per tree, each against the few
 * cylinders fitted last and a subset of the SAME cloud, keeping one bit per point (:1084).
 *
 * tm_cloud_upload_host: make an (n, >=3) TM_F32 / TM_F64 host cloud resident on the device as fp32 xyz
 *   (torch.tensor(points, dtype=torch.float32), :1076-1079 via Projection.py:33).  Synchronous.
 * tm_proximity_flags_host: for the rows `subset_host[0..n)` of the resident cloud (NULL: rows 0..n-1) and m RAW
 *   cylinders (start (m,3), end (m,3), radius (m,), fp32 host arrays): prepares the cylinders as :1043-1045 does
 *   (axis_eps 0) or as Projection.py:126-132 does (axis_eps 1e-8), runs closest_cylinder_cuda_batch with `params`
 *   (mode / cell_size ignored: every pair is evaluated) and returns, per row, out_flags = (distance < eps) (:1084),
 *   and optionally the distance and the winning cylinder row.  Any output may be NULL.  One H2D copy, one kernel,
 *   one D2H copy per output; synchronous.  m is limited to 3072; larger tables go through tm_set_cylinders.
 */
int tm_cloud_upload_host(tm_handle *h, const void *cloud_host, int32_t cloud_dtype, int64_t n,
                         int64_t cloud_row_stride);
int tm_proximity_flags_host(tm_handle *h, const int64_t *subset_host, int64_t n,
                            const float *start_host, const float *end_host, const float *radius_host, int64_t m,
                            const tm_params *params, float axis_eps, float eps,
                            uint8_t *out_flags_host, float *out_dist_host, int32_t *out_index_host);

/*
 * Point-neighbourhood features the drivers append right after the labelling (Modules/Features.py:178-229 at
 * LabelGenerationCuda.py:197-198 / Projection.py:420-427), device side.  pts: float64 DEVICE array, row i at
 * pts + i*row_stride, xyz contiguous (the labelled cloud is float64).
 *
 * tm_knn_covariance: for every point the k (2..32) nearest points of the same cloud, itself included (exact; ties by lower
 *   row), and np.cov of (neighbours - point) (mean-subtracted, / (k-1)) as a row-major 3x3 in out_cov (n,9) — the matrix
 *   compute_normals_ckdtree (:111-133, k = 15) feeds to the SVD and compute_curvature_ckdtree (:136-157, k = 10) to
 *   eigvalsh.  out_idx (n,k) rows of the neighbours in ascending distance, may be NULL.
 * tm_radius_count: len(tree.query_ball_point(point, r)) for every point (compute_density_ckdtree, :160-172).
 * Both return TM_ERR_INVALID for non-finite coordinates or n < k.
 */
int tm_knn_covariance(tm_handle *h, const double *pts, int64_t n, int64_t row_stride, int32_t k,
                      double *out_cov, int32_t *out_idx, void *stream);
int tm_radius_count(tm_handle *h, const double *pts, int64_t n, int64_t row_stride, double radius,
                    int32_t *out_count, void *stream);

/*
 * Noisy surface cloud of a QSM on the device (PreProcessing/NoiseDataGeneration.py:60-102; the producer of the clouds
 * the labelling path consumes).  The per-cylinder part of the reference (:33-58, :77-96: point counts from the
 * height-dependent density, Rodrigues rotations) is O(M) float64 numpy and stays with the caller, who passes
 *   cyl_rec      DEVICE (m,14) float64: start xyz, rotation matrix row-major (9), radius, axis length;
 *   first_point  DEVICE (m+1) int64: exclusive prefix sum of the per-cylinder point counts (first_point[m] = cloud size).
 * The call writes rows point0 .. point0+n of the cloud to out_points (DEVICE (n,3) float64; out_points_f32, optional, the
 * same rows rounded to float32 as Modules/Utils.py:236 load_cloud would hand them to the labeller).  Variates:
 *   theta/z/noise all NULL: drawn from Philox4x32-10, counter = (point number, block 0/1), key = seed — angle 2*pi*u,
 *     axial position L*u, radial noise exp(-3 + 0.85*g) with g = sqrt(-2 ln(1-u1)) cos(2*pi*u2) (np.random.uniform /
 *     np.random.lognormal(-3, 0.85) of :64-68 in distribution); a cloud depends on (table, seed) only, not on how the
 *     rows are split over calls or ranks;
 *   all three given (DEVICE (n,) float64, row i of this call): used as is — this is how the parity tests replay the
 *     reference's own Mersenne-Twister draws.
 * float64 arithmetic in the reference's order.  n == 0 is a no-op; m == 0 with n > 0 is TM_ERR_NO_CYLINDERS; rows at or
 * beyond first_point[m] come back as NaN.
 */
int tm_noise_cloud(tm_handle *h, const double *cyl_rec, const int64_t *first_point, int64_t m, int64_t n, int64_t point0,
                   uint64_t seed, const double *theta, const double *z, const double *noise,
                   double *out_points, float *out_points_f32, void *stream);

/*
 * How the last tm_label_cloud_host call moved its results: with >= 4 host threads available the (N,7) float64 records are
 * assembled by host worker threads (xyz from the caller's own cloud, 16 bytes of {offset, id} per point over PCIe;
 * *host_threads = workers used); otherwise the device assembles them and 56 bytes per point come back (*host_threads = 0).
 * TM_HOST_ASSEMBLE=0 in the environment forces the latter, =k forces k workers.
 */
int tm_host_pipeline_info(tm_handle *h, int32_t *d2h_bytes_per_point, int32_t *host_threads);

/*
 * STREAM-style probe of the HOST memory system with the worker threads tm_label_cloud_host uses: copies a 256 MiB
 * buffer (plain loads, streaming stores) a few times and reports (bytes read + bytes written) per second of the best
 * pass.  bench.py quotes the end-to-end call's host traffic against it.  No device work.
 */
int tm_measure_host_bandwidth(tm_handle *h, double *bytes_per_second, int32_t *threads);

/* Counters of the last labelling call.  Synchronises the device. */
int tm_get_stats(tm_handle *h, tm_stats *out);

/*
 * Per-phase device timing of tm_label_points (CUDA events recorded on the caller's stream between the
 * phases; off by default).  tm_get_phase_ms synchronises and fills out[0..TM_PHASES):
 *   [0] bin points into voxels   [1] voxel scan + work items   [2] scatter into voxel order
 *   [3] tile kernel: cull + dense evaluation, winning row -> original row   [4] ring / tree search of uncertified points
 *   [5] exhaustive kernel (brute mode, or the grid's outliers)   [6] winning rows of the pending points (brute mode: the
 *   winner epilogue)   [7] streaming winner epilogue: rows -> label + offset arrays   [8] whole call
 * Phases that did not run report 0.
 */
#define TM_PHASES 9
int tm_set_profiling(tm_handle *h, int enabled);
int tm_get_phase_ms(tm_handle *h, float *out_ms);

/*
 * FP32 FMA-pipe probe: dependent FFMA chains on every SM; reports lane-operations per second
 * (one FFMA = one lane-op, the convention of SURVEY.md A.6).  Used by bench.py as the FP32
 * roofline denominator because MEASURED_PEAKS.json has no FP32 entry.  Synchronous.
 */
int tm_measure_fp32_peak(tm_handle *h, double *lane_ops_per_second);

/*
 * Self-test of the division / square-root sequences of the pair evaluation (csrc/tm_eval.cuh: div3, sqrt_rn)
 * against the compiler's IEEE-rounded __fdiv_rn / __fsqrt_rn on n pseudo-random operand sets (the kernel's
 * regime, random bit patterns, edge values).  *mismatches must come back 0.  Synchronous.
 */
int tm_selftest_arithmetic(tm_handle *h, uint64_t n, uint32_t seed, uint64_t *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* TREEMORPH_NN_H */

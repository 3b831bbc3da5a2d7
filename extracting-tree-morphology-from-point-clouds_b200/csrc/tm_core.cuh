// Shared declarations of the library's translation units: the handle, scratch buffers, error
// plumbing and the internal entry points of the brute-force and voxel-grid paths.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/treemorph_nn.h"

namespace tmn {

// grow-only device scratch buffer owned by the handle
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const { return static_cast<T *>(p); }
};

// Page-locked host staging, grow-only.
struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap && p) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

constexpr int PIPE_SLOTS = 4;          // chunks in flight in tm_label_cloud_host

// Uniform voxel grid over the cylinders' solid AABBs (plus a margin).  Voxel (x,y,z) has the
// linear id morton-interleaved over (bits.x, bits.y, bits.z) so that neighbouring voxels are
// neighbours in memory and consecutive work items share candidate cylinders in L1/L2.
struct GridDesc {
    float ox, oy, oz;        // world position of the corner of voxel (0,0,0)
    float h, inv_h;          // voxel edge
    int nx, ny, nz;          // extent in voxels
    int bx, by, bz;          // bits per axis of the interleaved id
    uint32_t ncell_codes;    // 1 << (bx+by+bz): size of the per-voxel arrays
};

// device-side counters of one labelling call (mirrors tm_stats where it is data dependent)
struct DevStats {
    unsigned long long pairs_grid;   // full evaluations in the voxel-tile kernel
    unsigned long long pairs_ring;   // full evaluations in the tree-search and exhaustive-with-cull kernels
    unsigned long long cull_tests;   // capsule lower-bound tests, all kernels
    unsigned long long points_binned;
    unsigned int voxels_occupied;
    unsigned int work_items;         // occupied voxels with their tile descriptor (one item per voxel run)
    unsigned int pending;            // points the voxel-tile kernel could not certify (slots of pend_idx / keys)
    unsigned int n_brute;            // of those, points that need the exhaustive kernel (slots of brute_slots)
    unsigned int far_certified;      // points certified by the far part of their own tile (inside the tile kernel)
    unsigned int ring_certified;     // points certified by the ring search
    unsigned int bvh_cursor;         // next pending slot the tree search hands out
    unsigned int lane_slots;         // (voxel, <= 2 points) lane slots of the tile kernel
    unsigned int undecided_near;     // points the estimates left undecided but certified inside D_near (front of the list)
    unsigned int undecided_far;      // points the estimates could not certify inside D_near (back of the list)
    unsigned long long bound_tests;  // approximate-distance bounds computed by the tile kernel
    unsigned int late_rows;          // direct path: rows whose outputs are written by the list epilogue
};

// FP32 lane-operations of one closed-form distance estimate of the tile kernel (tm_grid.cu: bound_pair), counted the way
// SURVEY.md A.6 counts the reference's 81: every add / sub / mul / fma / min / max / compare / select / rsqrt is one
// (variant A as compiled: 10 FFMA + 4 FADD + 2 FMUL + 1 MUFU + 4 FSETP + 4 FMNMX + 1 FMNMX3 + 2 FSEL + 1 SEL per estimate in the
// SASS of the staged loop; variant B has one compare and one select more)
constexpr uint32_t LANE_OPS_PER_BOUND = 29;

struct HostPool;                     // tm_api.cu

}  // namespace tmn

struct tm_handle {
    int device = 0;
    int sm_count = 148;
    char err[512] = {0};

    // ---- cylinder table (tm_set_cylinders) ----
    int64_t m = 0;
    bool have_cyl = false;
    tmn::DevBuf recA, recB;          // float4[M]: {start.xyz, axis_length}, {unit.xyz, radius}
    tmn::DevBuf recAB;               // float4[2M]: the same records interleaved (one 32-byte sector per cylinder: gathers)
    tmn::DevBuf ids;                 // int32[M]
    tmn::DevBuf boxlo, boxhi;        // float4[M]: solid-cylinder AABB (w unused)
    tmn::DevBuf bbox;                // 6 floats as ordered ints: global min/max + counters + size statistics
    float mean_extent = 0.f;        // mean length of the regular cylinders (mean diameter if all lengths are zero)

    // ---- static voxel index of the cylinders (per table and cell size) ----
    bool have_grid = false;
    float grid_cell = 0.f;          // cell size the index was requested with
    float reach = 0.f;              // D_max: tile(V) holds every cylinder whose capsule comes within D_max of voxel V
    float near = 0.f;               // D_near: the leading `near` entries of a tile are those within D_near of the voxel
    float maxabs = 0.f;             // largest |coordinate| of the grid (scales the rounding slack)
    float slack_floor = 1e-4f;      // absolute part of the rounding slack (proportional to the voxel edge below tree scale)
    tmn::GridDesc grid{};
    tmn::DevBuf cyl_cell_start;      // uint32[ncell_codes + 1]: first pool entry of each voxel's tile (multiple of 4)
    tmn::DevBuf cyl_cell_cnt;        // uint32[ncell_codes]: tile length
    tmn::DevBuf cyl_cell_near;       // uint32[ncell_codes]: length of the tile's near part (lower bound <= D_near)
    tmn::DevBuf tileAB, tileI;       // tile pool: interleaved records {A, B} (one 32-byte sector) + cylinder row of every
    tmn::DevBuf tileLB;              //            (voxel, cylinder) entry, sorted per voxel by tileLB = lower bound of dist(voxel box, capsule)
    tmn::DevBuf tile_keys;           // u64 per entry, build-time only: (bits(lower bound) << 32) | cylinder row
    tmn::DevBuf long_list;           // int32[]: cylinders whose dilated AABB spans too many voxels
    tmn::DevBuf special;             // int32[]: non-finite / non-unit cylinders, evaluated for every point
    tmn::DevBuf aligned;             // int32[]: axis-parallel cylinders (variant A: NaN for points on their axis LINE)
    uint32_t n_long = 0, n_special = 0, n_listed = 0, n_aligned = 0;
    tmn::DevBuf bvh_nodes;           // 64-byte nodes of the bounding-volume hierarchy over the regular cylinders
    tmn::DevBuf bvh_rows;            // int32: cylinder row per leaf slot (Morton order)
    tmn::DevBuf bvh_leafAB;          // float4[2 x count]: records in leaf order
    int32_t bvh_root = 0, bvh_count = 0;
    tmn::DevBuf bvh_scratch;         // build-time scratch of the BVH (sort keys, per-slot boxes, counters)
    uint64_t index_entries = 0;
    uint32_t voxels_with_tiles = 0;  // voxels within D_max of some cylinder (density estimate for the point sort)

    // ---- per-call scratch ----
    tmn::DevBuf keys;                // u64 per point (brute mode) / per pending slot (grid mode)
    tmn::DevBuf cell_count, cell_start, block_sums;      // index build (cell_count / cell_start) and scan partials
    tmn::DevBuf cells;               // uint2 per voxel: {point count -> scatter cursor, first sorted point}
    tmn::DevBuf sorted_pts;          // float4 per point {x,y,z,bits(original row)}
    tmn::DevBuf items;               // 2 x uint4 per occupied voxel {tile offset, near length, first sorted point, point count} {far length, first lane slot, -, -}
    tmn::DevBuf warp_item;           // uint32 per group of 32 lane slots: the item its first slot belongs to
    tmn::DevBuf undecided;           // uint4 per point the estimates left undecided (exact kernel's work list)
    tmn::DevBuf late_rows;           // uint32: rows the direct kernel did not finish itself
    tmn::DevBuf tile_desc;           // uint4 per voxel code {tile offset, near length, tile length, 0} (static, per table)
    tmn::DevBuf pend_idx;            // int32 original row per pending slot (sign bit: outside the grid)
    tmn::DevBuf brute_slots;         // uint32 pending slots that need the exhaustive kernel
    tmn::DevBuf win;                 // int32 per point: winning cylinder row (when the caller passes no out_index)
    tmn::DevBuf dstats;              // tmn::DevStats + cursors
    tmn::DevBuf scratch_f;           // misc float scratch

    // ---- host pipeline (tm_label_cloud_host) ----
    cudaStream_t pipe_stream[4] = {nullptr, nullptr, nullptr, nullptr};      // H2D, label, D2H, D2H of device-assembled rows
    cudaEvent_t pipe_event[4 * tmn::PIPE_SLOTS] = {nullptr};
    tmn::PinnedBuf pinned_in[tmn::PIPE_SLOTS], pinned_out[tmn::PIPE_SLOTS];
    tmn::DevBuf chunk_in[tmn::PIPE_SLOTS], chunk_rec[tmn::PIPE_SLOTS], chunk_off[tmn::PIPE_SLOTS], chunk_id[tmn::PIPE_SLOTS],
        chunk_dist[tmn::PIPE_SLOTS];

    // ---- small-table fast path (tm_cloud_upload_host / tm_proximity_flags_host) ----
    tmn::DevBuf cloud_res;           // resident (n,3) fp32 copy of the caller's cloud
    int64_t cloud_res_n = -1;
    tmn::DevBuf small_in;            // per call: subset rows (int64) | cylinders (7 floats each)
    tmn::DevBuf small_out;           // per call: flags (u8) | dist (f32) | index (i32)
    cudaStream_t small_stream = nullptr;

    tmn::HostPool *pool = nullptr;   // host worker threads that assemble the (N,7) float64 records
    int32_t host_d2h_bytes_per_point = 0, host_assembly_threads = 0;     // what the last tm_label_cloud_host call did
    tmn::DevBuf chunk_packed[tmn::PIPE_SLOTS];

    // ---- point features (tm_knn.cu) ----
    tmn::DevBuf knn_cells, knn_start, knn_sorted, knn_box;

    // ---- overlapped epilogue (tm_grid.cu) ----
    cudaStream_t side_stream = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;

    // ---- multi-GPU (tm_comm.cu) ----
    void *comm = nullptr;            // ncclComm_t
    int32_t comm_rank = 0, comm_size = 0;
    tmn::DevBuf comm_table, comm_count;      // packed (M,9) table as broadcast, its row count

    // ---- optional phase timing ----
    bool profiling = false;
    cudaEvent_t phase_ev[TM_PHASES + 1] = {nullptr};
    bool phase_hit[TM_PHASES + 1] = {false};

    tm_stats stats{};
    int64_t last_n = 0;             // point count of the most recent labelling call (for tm_get_stats)
};

namespace tmn {

inline int fail(tm_handle *h, int code, const char *fmt, const char *a = "", const char *b = "") {
    if (h) snprintf(h->err, sizeof(h->err), fmt, a, b);
    return code;
}

#define TM_CUDA(h, expr)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            return tmn::fail((h), _e == cudaErrorMemoryAllocation ? TM_ERR_NOMEM : TM_ERR_CUDA,   \
                            "%s failed: %s", #expr, cudaGetErrorString(_e));                     \
        }                                                                                        \
    } while (0)

// after a kernel launch: launch errors always; with TM_DEBUG_SYNC=1 in the environment also execution errors,
// attributed to the kernel that raised them
#define TM_KCHECK(h, st, name)                                                                   \
    do {                                                                                         \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e == cudaSuccess && tmn::debug_sync()) _e = cudaStreamSynchronize(st);              \
        if (_e != cudaSuccess) return tmn::fail((h), TM_ERR_CUDA, "%s: %s", name, cudaGetErrorString(_e)); \
    } while (0)

inline bool debug_sync() {
    static const bool on = [] { const char *e = getenv("TM_DEBUG_SYNC"); return e && e[0] == '1'; }();
    return on;
}

// phase marks: event i is recorded when phase i-1 ends / phase i begins (0 = call start, TM_PHASES = call end)
inline void mark(tm_handle *h, int i, cudaStream_t st) {
    if (!h->profiling) return;
    if (!h->phase_ev[i]) cudaEventCreate(&h->phase_ev[i]);
    cudaEventRecord(h->phase_ev[i], st);
    h->phase_hit[i] = true;
}

struct LabelArgs {
    const float *pts;
    int64_t n;
    int64_t row_stride;
    tm_params prm;
    int32_t *out_index;
    int32_t *out_id;
    float *out_dist;
    float *out_offset;
    float *out_radius;
    cudaStream_t stream;
    float4 *out_packed = nullptr;    // optional {offset.xyz, bits(id)} per row (the host-assembly path of tm_label_cloud_host)
};

// tm_brute.cu
int label_brute(tm_handle *h, const LabelArgs &a);
// grid mode, after the voxel-tile and tree-search kernels: exhaustive search (with the capsule cull) for the pending
// slots listed in h->brute_slots, then the winning row of EVERY pending slot goes to win[original row]
int finish_pending(tm_handle *h, const LabelArgs &a, DevStats *dst, int32_t *win, float maxabs);
// streaming winner-only epilogue over all rows: win[row] -> index / id / distance / offset / radius
// rows == nullptr: every row 0..n-1; else the *d_count rows listed in `rows`
int finalize_rows(tm_handle *h, const LabelArgs &a, const int32_t *win, const uint32_t *rows, const unsigned int *d_count);
// tm_small.cu
constexpr int SMALL_MAX_M = 3072;     // cylinders per call of the small-table kernel (2 x 16 B of shared memory each)
struct SmallArgs {
    const float *cloud;            // resident (n_cloud, 3) fp32
    const int64_t *subset;         // rows to process (nullptr: rows 0..n-1)
    int64_t n;
    const float *cyl;              // m rows of 7 floats: start xyz, end xyz, radius
    int m;
    float axis_eps, atol, eps_norm, eps_flag;
    uint8_t *flags;
    float *dist;
    int32_t *index;
};
int run_proximity(tm_handle *h, const SmallArgs &a, bool guard, bool nfma, cudaStream_t st);
int exclusive_scan_u32(tm_handle *h, const uint32_t *count, uint32_t n, uint32_t *start, cudaStream_t stream);
// tm_bvh.cu
int build_bvh(tm_handle *h, cudaStream_t stream, int n_regular, const float *lo, const float *hi);
int search_bvh(tm_handle *h, const LabelArgs &a, DevStats *dst, int32_t *win);
// tm_grid.cu
int build_cylinder_index(tm_handle *h, float cell_size, cudaStream_t stream);
int label_grid(tm_handle *h, const LabelArgs &a);
float auto_cell_size(const tm_handle *h, int64_t n);

}  // namespace tmn

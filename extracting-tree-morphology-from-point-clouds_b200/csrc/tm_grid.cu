// Voxel-grid path: exact nearest-cylinder search with spatial pruning.
//
//   per table (tm_set_cylinders + first use of a cell size):
//     every voxel V gets a TILE: the cylinders whose capsule comes within D_max of box(V), sorted by
//     lb(V, c) = a lower bound of dist(box(V), capsule(c)), packed contiguously (float4 A | float4 B | row | lb).
//     The leading `near(V)` entries are those with lb <= D_near.
//   per call:
//     points --count/scan/scatter--> counting sort by voxel id (brick-Morton order) --> contiguous per-voxel runs
//     evaluate:  one warp per (voxel, <= 64 points) work item.  The NEAR part of the voxel's tile is staged into shared
//                memory with bulk async copies (double buffered, next item's tile prefetched).  For every tile entry
//                each lane runs a 17-instruction capsule lower-bound test for its (up to) two points against the
//                point's incumbent; surviving (point, entry) pairs are compacted into a per-warp queue (ballot +
//                popc) and evaluated 32 at a time with the reference arithmetic, so the expensive evaluation always
//                runs with full lanes.  Winners are merged with a 64-bit (distance, index) atomicMin in shared
//                memory, which is torch.argmin's comparator.
//                A point whose best distance is <= D_near is CERTIFIED: every cylinder that could beat or tie it has
//                lb <= D_near for the point's voxel, i.e. sits in the near part.  The winning row is stored at the
//                point's original row (4 bytes, L2 resident).
//                The few points that are not certified (noise tail) then walk the FAR part of the tile, lanes across
//                entries in ascending lb order, and stop at the first entry whose lb exceeds their incumbent.
//     ring:      a handful of points still uncertified at D_max (noise tail): ball query over the neighbouring voxels'
//                tiles, one CTA per point (latency-optimised).
//     tree:      MANY such points (clutter far from every cylinder) and points outside the grid descend the
//                bounding-volume hierarchy of the cylinders (tm_bvh.cu), one thread per point (throughput-optimised).
//     brute:     non-finite points: exhaustive search (nothing bounds them).
//     epilogue:  streaming pass over the rows in input order: recompute the winning pair with full geometry, move to the
//                mantle, gather the ID, write label + offset (coalesced reads and writes).
//
// Exactness: dist_ref(p,c) >= dist(p, capsule(c)) >= lb(V,c) for p in V; the cull, the lb order and
// the certification only ever discard cylinders whose capsule is farther than the incumbent (plus a rounding
// allowance), so the argmin and its lowest-index tie-break are those of the exhaustive search.  Cylinders that cannot
// be bounded (non-finite, non-unit axis) are evaluated for every point; axis-parallel cylinders get the exact
// on-axis-line test in variant A (NaN wins the argmin at any distance).
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdlib>

#include "tm_core.cuh"
#include "tm_eval.cuh"
#include "tm_ptx.cuh"

namespace tmn {

// ------------------------------------------------------------------------------------------------
// voxel ids: linear over 8x8x8 bricks, 3-D Morton inside a brick
// ------------------------------------------------------------------------------------------------
struct GridDev {
    float ox, oy, oz, h, inv_h;
    int nx, ny, nz;        // voxels
    int bnx, bny, bnz;     // bricks
    float slack;           // fp32 rounding allowance of the reference pipeline at this coordinate scale
    float reach;           // D_max: radius covered by a whole tile
    float near;            // D_near: radius covered by the near part of a tile
};

__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {      // 3 bits -> every third bit
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4);
}
__host__ __device__ __forceinline__ uint32_t voxel_code(const GridDev &g, int x, int y, int z) {
    const uint32_t brick = (static_cast<uint32_t>(z >> 3) * g.bny + static_cast<uint32_t>(y >> 3)) * g.bnx +
                           static_cast<uint32_t>(x >> 3);
    return (brick << 9) | spread3(x & 7) | (spread3(y & 7) << 1) | (spread3(z & 7) << 2);
}

static GridDev to_dev(const GridDesc &d, float slack, float reach, float near) {
    GridDev g;
    g.ox = d.ox; g.oy = d.oy; g.oz = d.oz; g.h = d.h; g.inv_h = d.inv_h;
    g.nx = d.nx; g.ny = d.ny; g.nz = d.nz;
    g.bnx = (d.nx + 7) / 8; g.bny = (d.ny + 7) / 8; g.bnz = (d.nz + 7) / 8;
    g.slack = slack;
    g.reach = reach;
    g.near = near;
    return g;
}

__device__ __forceinline__ int cell_coord(float p, float o, float inv_h, int n) {
    const float f = (p - o) * inv_h;
    int c = static_cast<int>(floorf(f));
    return min(max(c, 0), n - 1);
}

// ------------------------------------------------------------------------------------------------
// cylinder side: static tiles
// ------------------------------------------------------------------------------------------------
constexpr int LONG_CELLS = 1 << 15;   // dilated AABBs spanning more voxels than this go to the "long" list

// dist(axis segment of the cylinder, box [lo, lo + h]^3), from below.  f(t) = dist^2(start + t * unit, box) is convex
// in t, so a golden-section search keeps the minimiser bracketed; the distance is 1-Lipschitz in t (|unit| = 1), hence
// min over the final bracket >= sqrt(best sample) - bracket width.
__device__ __forceinline__ float seg_box_dist2(const float4 A, const float4 B, float lx, float ly, float lz, float h, float t) {
    const float x = fmaf(t, B.x, A.x), y = fmaf(t, B.y, A.y), z = fmaf(t, B.z, A.z);
    const float gx = fmaxf(fmaxf(lx - x, x - (lx + h)), 0.f);
    const float gy = fmaxf(fmaxf(ly - y, y - (ly + h)), 0.f);
    const float gz = fmaxf(fmaxf(lz - z, z - (lz + h)), 0.f);
    return fmaf(gz, gz, fmaf(gy, gy, gx * gx));
}

__device__ __forceinline__ float seg_box_lower_bound(const float4 A, const float4 B, float lx, float ly, float lz, float h) {
    constexpr float INVPHI = 0.61803398875f;
    float lo = 0.f, hi = A.w;
    float c = hi - (hi - lo) * INVPHI, d = lo + (hi - lo) * INVPHI;
    float fc = seg_box_dist2(A, B, lx, ly, lz, h, c), fd = seg_box_dist2(A, B, lx, ly, lz, h, d);
    float best = fminf(fminf(fc, fd), fminf(seg_box_dist2(A, B, lx, ly, lz, h, lo), seg_box_dist2(A, B, lx, ly, lz, h, hi)));
#pragma unroll 1
    for (int it = 0; it < 26; ++it) {
        if (fc < fd) { hi = d; d = c; fd = fc; c = hi - (hi - lo) * INVPHI; fc = seg_box_dist2(A, B, lx, ly, lz, h, c); best = fminf(best, fc); }
        else         { lo = c; c = d; fc = fd; d = lo + (hi - lo) * INVPHI; fd = seg_box_dist2(A, B, lx, ly, lz, h, d); best = fminf(best, fd); }
    }
    return sqrtf(best) - (hi - lo);
}

// One warp per cylinder.  pass 0 counts, pass 1 fills: voxel V receives cylinder c when
//   lb_raw(V, c) = dist(box(V), axis segment) - |r|  <=  D_max + slack,
// and the entry carries lb = max(0, lb_raw - slack - margin), a lower bound of the reference distance between any
// point binned into V and c (slack absorbs the fp32 rounding of the binning, of this arithmetic and of the
// reference pipeline).  Entries are written as sortable keys (bits(lb) << 32 | c); tile_sort_kernel orders them.
__global__ void __launch_bounds__(256)
cyl_register_kernel(const float4 *__restrict__ recA, const float4 *__restrict__ recB, const float4 *__restrict__ boxlo,
                    const float4 *__restrict__ boxhi, int m, GridDev g, int pass, uint32_t *__restrict__ cell_counter,
                    const uint32_t *__restrict__ cell_start, unsigned long long *__restrict__ tile_keys,
                    int32_t *__restrict__ long_list, unsigned int *__restrict__ n_long) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= m) return;
    const float4 lo = boxlo[c], hi = boxhi[c];
    if (!(lo.w == 0.f)) return;                       // w != 0 marks special (non-finite / non-unit) cylinders
    const float grow = g.reach + g.slack;
    const int x0 = cell_coord(lo.x - grow, g.ox, g.inv_h, g.nx), x1 = cell_coord(hi.x + grow, g.ox, g.inv_h, g.nx);
    const int y0 = cell_coord(lo.y - grow, g.oy, g.inv_h, g.ny), y1 = cell_coord(hi.y + grow, g.oy, g.inv_h, g.ny);
    const int z0 = cell_coord(lo.z - grow, g.oz, g.inv_h, g.nz), z1 = cell_coord(hi.z + grow, g.oz, g.inv_h, g.nz);
    const int sx = x1 - x0 + 1, sy = y1 - y0 + 1, sz = z1 - z0 + 1;
    const long long cells = static_cast<long long>(sx) * sy * sz;
    if (cells > LONG_CELLS) {
        if (pass == 0 && lane == 0) { const unsigned int s = atomicAdd(n_long, 1u); long_list[s] = c; }
        return;
    }
    const float4 A = recA[c], B = recB[c];
    const float ar = fabsf(B.w);
    const float lim = grow + 0.8660254f * g.h * 1.0001f + ar;      // cheap reject: capsule vs the voxel's circumsphere
    const float lim2 = lim * lim;
    for (int idx = lane; idx < static_cast<int>(cells); idx += 32) {
        const int x = x0 + idx % sx, y = y0 + (idx / sx) % sy, z = z0 + idx / (sx * sy);
        const float lx = g.ox + x * g.h, ly = g.oy + y * g.h, lz = g.oz + z * g.h;
        const float vx = lx + 0.5f * g.h - A.x, vy = ly + 0.5f * g.h - A.y, vz = lz + 0.5f * g.h - A.z;
        const float t = fminf(fmaxf(vx * B.x + vy * B.y + vz * B.z, 0.f), A.w);
        const float wx = vx - t * B.x, wy = vy - t * B.y, wz = vz - t * B.z;
        if (wx * wx + wy * wy + wz * wz > lim2) continue;
        const float lb_raw = seg_box_lower_bound(A, B, lx, ly, lz, g.h) - ar;
        if (lb_raw > grow) continue;
        const uint32_t code = voxel_code(g, x, y, z);
        const uint32_t s = atomicAdd(&cell_counter[code], 1u);
        if (pass == 1) {
            const float lb = fmaxf(lb_raw - 2.f * g.slack, 0.f);
            tile_keys[cell_start[code] + s] = (static_cast<unsigned long long>(__float_as_uint(lb)) << 32) | static_cast<uint32_t>(c);
        }
    }
}

// One warp per voxel: order the tile's keys by (lb, cylinder row) and expand them into the pool arrays.
// Tiles longer than SORT_MAX keep their arbitrary order with lb = 0 (always a valid lower bound): everything is "near".
constexpr int SORT_MAX = 1024;
constexpr int SORT_WARPS = 8;

__global__ void __launch_bounds__(SORT_WARPS * 32)
tile_sort_kernel(const uint32_t *__restrict__ cell_start, const uint32_t *__restrict__ cell_cnt, uint32_t ncodes,
                 const unsigned long long *__restrict__ tile_keys, const float4 *__restrict__ recA,
                 const float4 *__restrict__ recB, float near_reach, float4 *__restrict__ tileA, float4 *__restrict__ tileB,
                 int32_t *__restrict__ tileI, float *__restrict__ tileLB, uint32_t *__restrict__ cell_near,
                 unsigned int *__restrict__ n_with_tiles) {
    extern __shared__ __align__(16) unsigned char sort_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long *buf = reinterpret_cast<unsigned long long *>(sort_smem) + static_cast<size_t>(warp) * SORT_MAX;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned int with_tiles = 0;
    for (uint32_t code = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; code < ncodes; code += nwarps) {
        const uint32_t n = cell_cnt[code];
        if (n == 0) { if (lane == 0) cell_near[code] = 0; continue; }
        ++with_tiles;
        const uint32_t off = cell_start[code];
        uint32_t near = 0;
        if (n <= SORT_MAX) {
            uint32_t P = 32;
            while (P < n) P <<= 1;
            for (uint32_t i = lane; i < P; i += 32) buf[i] = i < n ? tile_keys[off + i] : KEY_NONE;
            __syncwarp();
            for (uint32_t k = 2; k <= P; k <<= 1) {
                for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                    for (uint32_t i = lane; i < P; i += 32) {
                        const uint32_t ixj = i ^ j;
                        if (ixj > i) {
                            const unsigned long long a = buf[i], b = buf[ixj];
                            const bool asc = (i & k) == 0;
                            if ((a > b) == asc) { buf[i] = b; buf[ixj] = a; }
                        }
                    }
                    __syncwarp();
                }
            }
            for (uint32_t i = lane; i < n; i += 32) {
                const unsigned long long key = buf[i];
                const uint32_t c = static_cast<uint32_t>(key);
                const float lb = __uint_as_float(static_cast<uint32_t>(key >> 32));
                tileA[off + i] = recA[c];
                tileB[off + i] = recB[c];
                tileI[off + i] = static_cast<int32_t>(c);
                tileLB[off + i] = lb;
                near += lb <= near_reach ? 1u : 0u;
            }
            __syncwarp();
        } else {
            for (uint32_t i = lane; i < n; i += 32) {
                const uint32_t c = static_cast<uint32_t>(tile_keys[off + i]);
                tileA[off + i] = recA[c];
                tileB[off + i] = recB[c];
                tileI[off + i] = static_cast<int32_t>(c);
                tileLB[off + i] = 0.f;
            }
            near = lane == 0 ? n : 0u;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) near += __shfl_xor_sync(0xffffffffu, near, o);
        if (n - near > 0xFFFFFFu) near = n;          // the far length must fit the 24 bits of a work item
        if (lane == 0) cell_near[code] = near;
    }
    if (lane == 0 && with_tiles) atomicAdd(n_with_tiles, with_tiles);
}

// tile lengths rounded up to 4 entries so every tile starts 16-byte aligned in all three pool arrays
__global__ void align4_kernel(const uint32_t *__restrict__ cnt, uint32_t *__restrict__ cnt_keep, uint32_t *__restrict__ rounded,
                              uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cnt[i];
    cnt_keep[i] = c;
    rounded[i] = (c + 3u) & ~3u;
}

// ------------------------------------------------------------------------------------------------
// generic 3-channel exclusive scan over the voxel arrays (points, occupied flags, work items)
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;     // 4096 voxels per block
constexpr int PTS_PER_ITEM = 64;
constexpr int CELL_PAD = 4;               // uint2 slots per voxel cell: one 32-byte sector each, so that neighbouring voxels'
                                          // atomics do not queue up on a shared sector

struct Tri { uint32_t a, b, c; };
__device__ __forceinline__ Tri tri_add(Tri x, Tri y) { return Tri{x.a + y.a, x.b + y.b, x.c + y.c}; }

__device__ __forceinline__ Tri block_exclusive_scan(Tri v, Tri *total) {
    __shared__ Tri warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Tri inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Tri n;
        n.a = __shfl_up_sync(0xffffffffu, inc.a, o);
        n.b = __shfl_up_sync(0xffffffffu, inc.b, o);
        n.c = __shfl_up_sync(0xffffffffu, inc.c, o);
        if (lane >= o) inc = tri_add(inc, n);
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        Tri w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : Tri{0, 0, 0};
        Tri winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            Tri n;
            n.a = __shfl_up_sync(0xffffffffu, winc.a, o);
            n.b = __shfl_up_sync(0xffffffffu, winc.b, o);
            n.c = __shfl_up_sync(0xffffffffu, winc.c, o);
            if (lane >= o) winc = tri_add(winc, n);
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = Tri{winc.a - w.a, winc.b - w.b, winc.c - w.c};
        if (lane == SCAN_THREADS / 32 - 1 && total) *total = winc;
    }
    __syncthreads();
    const Tri base = warp_sums[warp];
    Tri ex = Tri{inc.a - v.a + base.a, inc.b - v.b + base.b, inc.c - v.c + base.c};
    __syncthreads();
    return ex;
}

__device__ __forceinline__ Tri tri_of_count(uint32_t cnt) {
    return Tri{cnt, cnt ? 1u : 0u, (cnt + PTS_PER_ITEM - 1) / PTS_PER_ITEM};
}

// phase A: per-block totals
// A voxel's counter may be split into `nsub` sub-cells (dense clouds: the atomics of one voxel spread over several
// addresses); the voxel's count is their sum, its run is the concatenation of the sub-runs.
template <int NSUB>
__device__ __forceinline__ uint32_t voxel_count(const uint32_t *count, int stride, uint32_t code) {
    uint32_t c[NSUB];
#pragma unroll
    for (int sidx = 0; sidx < NSUB; ++sidx) c[sidx] = count[(static_cast<size_t>(code) * NSUB + sidx) * stride];   // loads in flight together
    uint32_t sum = 0;
#pragma unroll
    for (int sidx = 0; sidx < NSUB; ++sidx) sum += c[sidx];
    return sum;
}

template <int NSUB>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t *__restrict__ count, int stride, uint32_t ncodes,
                                                                   Tri *__restrict__ block_sums) {
    __shared__ Tri total;
    const uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    Tri v{0, 0, 0};
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < ncodes) v = tri_add(v, tri_of_count(voxel_count<NSUB>(count, stride, base + i)));
    block_exclusive_scan(v, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// phase B: one block scans the block totals in place (exclusive) and publishes the grand totals
__global__ void __launch_bounds__(SCAN_THREADS) scan_blocks_kernel(Tri *__restrict__ block_sums, uint32_t nblocks,
                                                                   DevStats *__restrict__ st) {
    __shared__ Tri total;
    __shared__ Tri carry;
    if (threadIdx.x == 0) carry = Tri{0, 0, 0};
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += SCAN_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const Tri v = i < nblocks ? block_sums[i] : Tri{0, 0, 0};
        const Tri ex = block_exclusive_scan(v, &total);
        const Tri c = carry;
        if (i < nblocks) block_sums[i] = tri_add(ex, c);
        __syncthreads();
        if (threadIdx.x == 0) carry = tri_add(c, total);
        __syncthreads();
    }
    if (threadIdx.x == 0 && st) {
        st->points_binned = carry.a;
        st->voxels_occupied = carry.b;
        st->work_items = carry.c;
    }
}

// phase C: final offsets.  mode 0 (tile pool): start[code] only (+ the grand total at start[ncodes]).
// In mode 1 `count` and `start` are the SAME array of {count, start} cells (stride 2): a thread reads the counts of its
// own cells, then rewrites them as {0, start} — the low word becomes the scatter pass's cursor, so that ONE 64-bit
// atomicAdd returns both the run start and the slot inside the run.
// mode 1 (points): start[code] = first sorted point of the voxel, and the voxel's work items
// {tile offset, near length, first point, point count <= 64 | far length << 8} are emitted in voxel-id order.
template <int NSUB>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t *count, int stride, uint32_t ncodes,
                                                                  const Tri *__restrict__ block_sums, int mode,
                                                                  uint32_t *start,
                                                                  const uint32_t *__restrict__ tile_start,
                                                                  const uint32_t *__restrict__ tile_cnt,
                                                                  const uint32_t *__restrict__ tile_near,
                                                                  uint4 *__restrict__ items) {
    const int lane = threadIdx.x & 31;
    const uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    uint32_t cnt[SCAN_ITEMS];
    Tri v{0, 0, 0};
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        cnt[i] = (base + i < ncodes) ? voxel_count<NSUB>(count, stride, base + i) : 0u;
        v = tri_add(v, tri_of_count(cnt[i]));
    }
    Tri run = tri_add(block_exclusive_scan(v, nullptr), block_sums[blockIdx.x]);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        uint32_t toff = 0, tnear = 0, far = 0, n_items = 0;
        if (base + i < ncodes) {
            if (mode == 0) {
                start[base + i] = run.a;
            } else {
                uint32_t c[NSUB];
#pragma unroll
                for (int sidx = 0; sidx < NSUB; ++sidx) c[sidx] = count[(static_cast<size_t>(base + i) * NSUB + sidx) * stride];
                uint32_t sub_start = run.a;
#pragma unroll
                for (int sidx = 0; sidx < NSUB; ++sidx) {
                    *reinterpret_cast<uint2 *>(start + (static_cast<size_t>(base + i) * NSUB + sidx) * stride) = make_uint2(0u, sub_start);
                    sub_start += c[sidx];
                }
                if (cnt[i]) {
                    toff = tile_start[base + i];
                    tnear = tile_near[base + i];
                    far = tile_cnt[base + i] - tnear;                // < 2^24 (tile_sort_kernel)
                    n_items = (cnt[i] + PTS_PER_ITEM - 1) / PTS_PER_ITEM;
                }
            }
        }
        // work items of the voxel: a few -> this thread; a crowded voxel (dense clouds: hundreds of items) -> the whole
        // warp, so that no thread ends up writing thousands of items alone
        if (n_items && n_items <= 4)
            for (uint32_t t = 0; t < n_items; ++t)
                items[run.c + t] = make_uint4(toff, tnear, run.a + t * PTS_PER_ITEM,
                                              min(static_cast<uint32_t>(PTS_PER_ITEM), cnt[i] - t * PTS_PER_ITEM) | (far << 8));
        uint32_t crowded = __ballot_sync(0xffffffffu, n_items > 4);
        while (crowded) {
            const int src = __ffs(crowded) - 1;
            crowded &= crowded - 1;
            const uint32_t s_toff = __shfl_sync(0xffffffffu, toff, src), s_near = __shfl_sync(0xffffffffu, tnear, src),
                           s_far = __shfl_sync(0xffffffffu, far, src), s_first = __shfl_sync(0xffffffffu, run.c, src),
                           s_pt = __shfl_sync(0xffffffffu, run.a, src), s_cnt = __shfl_sync(0xffffffffu, cnt[i], src);
            const uint32_t s_items = (s_cnt + PTS_PER_ITEM - 1) / PTS_PER_ITEM;
            for (uint32_t t = lane; t < s_items; t += 32)
                items[s_first + t] = make_uint4(s_toff, s_near, s_pt + t * PTS_PER_ITEM,
                                                min(static_cast<uint32_t>(PTS_PER_ITEM), s_cnt - t * PTS_PER_ITEM) | (s_far << 8));
        }
        run = tri_add(run, tri_of_count(cnt[i]));
    }
    if (mode == 0 && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) start[ncodes] = run.a;
}

static int run_scan(tm_handle *h, const uint32_t *count, uint32_t ncodes, int mode, uint32_t *start, const uint32_t *tile_start,
                    const uint32_t *tile_cnt, const uint32_t *tile_near, uint4 *items, DevStats *st, cudaStream_t stream, int nsub = 1) {
    const uint32_t nblocks = (ncodes + SCAN_BLOCK - 1) / SCAN_BLOCK;
    TM_CUDA(h, h->block_sums.ensure(sizeof(Tri) * nblocks));
    Tri *bs = h->block_sums.as<Tri>();
    const int stride = mode == 1 ? 2 * CELL_PAD : 1;
#define TM_SCAN_CASE(S)                                                                                                        \
    do {                                                                                                                       \
        scan_reduce_kernel<S><<<nblocks, SCAN_THREADS, 0, stream>>>(count, stride, ncodes, bs);                                \
        scan_blocks_kernel<<<1, SCAN_THREADS, 0, stream>>>(bs, nblocks, st);                                                   \
        scan_apply_kernel<S><<<nblocks, SCAN_THREADS, 0, stream>>>(count, stride, ncodes, bs, mode, start, tile_start, tile_cnt, \
                                                                   tile_near, items);                                          \
    } while (0)
    switch (nsub) {
        case 2: TM_SCAN_CASE(2); break;
        case 4: TM_SCAN_CASE(4); break;
        case 8: TM_SCAN_CASE(8); break;
        default: TM_SCAN_CASE(1); break;
    }
#undef TM_SCAN_CASE
    TM_CUDA(h, cudaGetLastError());
    return TM_OK;
}

// exclusive scan of `n` counters into start[0..n] (start[n] = total); used by the point-feature kernels (tm_knn.cu)
int exclusive_scan_u32(tm_handle *h, const uint32_t *count, uint32_t n, uint32_t *start, cudaStream_t stream) {
    return run_scan(h, count, n, 0, start, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}

// ------------------------------------------------------------------------------------------------
// host: size the grid from the cylinders' bounding box and build the static tiles
// ------------------------------------------------------------------------------------------------
static inline float ordered_to_float(int k) {
    int i = k >= 0 ? k : k ^ 0x7fffffff;
    float f;
    memcpy(&f, &i, 4);
    return f;
}

constexpr uint32_t MAX_CODES = 1u << 26;

// Voxel edge when the caller does not choose one: 1.43 x the mean cylinder length, i.e. a voxel holds a handful of
// cylinder pieces whatever the units or the scale of the model (0.25 m for tree QSMs with their ~0.175 m cylinders: the
// measured optimum is flat between 0.2 and 0.3 m).  Two significant bits of mantissa are kept
// so that the edge — and with it the static index — does not change with the last digits of the statistics.
float auto_cell_size(const tm_handle *h, int64_t) {
    float e = h->mean_extent;
    if (!(e > 0.f) || !std::isfinite(e)) return 0.25f;
    e = std::min(std::max(1.43f * e, 1e-5f), 1e5f);
    int ex;
    const float mant = std::frexp(e, &ex);                 // e = mant * 2^ex, mant in [0.5, 1)
    return std::ldexp(std::round(mant * 8.f) / 8.f, ex);
}

// rounding allowance of the reference's fp32 pipeline (and of the bound arithmetic) at coordinate scale `maxabs`
// The relative term covers ~30 ulp of the largest coordinate in play; the floor (0.1 mm at tree scale, proportional to the
// voxel edge for smaller models) is head-room on top of it.
static inline float slack_floor_for(float h) { return std::min(1e-4f, 4e-4f * h); }
static inline float slack_for(float maxabs, float h) { return slack_floor_for(h) + 4e-6f * maxabs; }

int build_cylinder_index(tm_handle *h, float cell_size, cudaStream_t stream) {
    static const bool trace_build = [] { const char *e = getenv("TM_TRACE_BUILD"); return e && e[0] == '1'; }();
    const auto t_begin = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace_build) return;
        cudaStreamSynchronize(stream);
        fprintf(stderr, "[tm build] %-22s %8.3f ms\n", what,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    // global bounding box of the regular cylinders' AABBs (written by pack_kernel as ordered ints)
    int host_box[10];
    TM_CUDA(h, cudaMemcpyAsync(host_box, h->bbox.p, sizeof(host_box), cudaMemcpyDeviceToHost, stream));
    TM_CUDA(h, cudaStreamSynchronize(stream));
    h->n_special = static_cast<uint32_t>(host_box[6]);
    const uint32_t n_regular = static_cast<uint32_t>(host_box[7]);
    h->n_aligned = static_cast<uint32_t>(host_box[8]);
    h->have_grid = false;
    h->grid_cell = cell_size;
    h->n_long = 0;
    h->index_entries = 0;
    if (n_regular == 0) {               // nothing to index: every point goes to the exhaustive kernel
        h->grid = GridDesc{};
        h->have_grid = true;
        h->n_listed = 0;
        return TM_OK;
    }
    float lo[3], hi[3];
    for (int k = 0; k < 3; ++k) { lo[k] = ordered_to_float(host_box[k]); hi[k] = ordered_to_float(host_box[3 + k]); }
    float hcell = cell_size;
    GridDesc d{};
    float margin = 0.f;
    for (;;) {
        double codes = 1;
        int n[3];
        margin = 6.f * hcell;                       // points farther than this outside the QSM box go to the tree search
        for (int k = 0; k < 3; ++k) {
            n[k] = static_cast<int>(std::ceil((static_cast<double>(hi[k]) - lo[k] + 2.0 * margin) / hcell));
            n[k] = std::max(n[k], 1);
            codes *= ((n[k] + 7) / 8) * 8.0;
        }
        if (codes <= MAX_CODES) {
            d.ox = lo[0] - margin; d.oy = lo[1] - margin; d.oz = lo[2] - margin;
            d.h = hcell; d.inv_h = 1.0f / hcell;
            d.nx = n[0]; d.ny = n[1]; d.nz = n[2];
            d.ncell_codes = static_cast<uint32_t>(codes);
            break;
        }
        hcell *= 1.5f;                  // coarsen until the voxel arrays fit
    }
    d.bx = d.by = d.bz = 0;
    h->grid = d;
    // D_near = nfac * h covers the bulk of a surface-sampled cloud; D_max = dfac * h bounds the far part of the tiles
    // (tuning hooks for experiments; the defaults are the shipped values)
    float nfac = 0.8f, dfac = 2.0f;
    if (const char *env = getenv("TM_NEAR_FACTOR")) { const float v = static_cast<float>(atof(env)); if (v >= 0.1f && v <= 4.f) nfac = v; }
    if (const char *env = getenv("TM_REACH_FACTOR")) { const float v = static_cast<float>(atof(env)); if (v >= 0.25f && v <= 8.f) dfac = v; }
    if (dfac < nfac) dfac = nfac;
    h->near = nfac * hcell;
    h->reach = dfac * hcell;
    float maxabs = 0.f;
    for (int k = 0; k < 3; ++k) maxabs = std::max(maxabs, std::max(std::fabs(lo[k]), std::fabs(hi[k])) + margin);
    h->maxabs = maxabs;
    h->slack_floor = slack_floor_for(hcell);
    const GridDev g = to_dev(d, slack_for(maxabs, hcell), h->reach, h->near);

    const int m = static_cast<int>(h->m);
    const uint32_t ncodes = d.ncell_codes;
    TM_CUDA(h, h->cell_count.ensure(sizeof(uint32_t) * (static_cast<size_t>(ncodes) + 1)));
    TM_CUDA(h, h->cell_start.ensure(sizeof(uint32_t) * (static_cast<size_t>(ncodes) + 1)));
    TM_CUDA(h, h->cyl_cell_start.ensure(sizeof(uint32_t) * (static_cast<size_t>(ncodes) + 1)));
    TM_CUDA(h, h->cyl_cell_cnt.ensure(sizeof(uint32_t) * (static_cast<size_t>(ncodes) + 1)));
    TM_CUDA(h, h->cyl_cell_near.ensure(sizeof(uint32_t) * (static_cast<size_t>(ncodes) + 1)));
    TM_CUDA(h, h->long_list.ensure(sizeof(int32_t) * static_cast<size_t>(m)));
    TM_CUDA(h, h->dstats.ensure(sizeof(DevStats) + 64));
    uint32_t *counter = h->cell_count.as<uint32_t>();
    uint32_t *rounded = h->cell_start.as<uint32_t>();          // per-call scratch, free at this point
    unsigned int *d_nlong = reinterpret_cast<unsigned int *>(h->dstats.as<unsigned char>() + sizeof(DevStats));
    TM_CUDA(h, cudaMemsetAsync(counter, 0, sizeof(uint32_t) * ncodes, stream));
    TM_CUDA(h, cudaMemsetAsync(d_nlong, 0, sizeof(unsigned int), stream));
    const int blocks = (m + 7) / 8;                               // 8 warps per block, one warp per cylinder
    cyl_register_kernel<<<blocks, 256, 0, stream>>>(h->recA.as<float4>(), h->recB.as<float4>(), h->boxlo.as<float4>(),
                                                    h->boxhi.as<float4>(), m, g, 0, counter, nullptr, nullptr,
                                                    h->long_list.as<int32_t>(), d_nlong);
    TM_KCHECK(h, stream, "cyl_register_kernel (count)");
    lap("register (count)");
    align4_kernel<<<(ncodes + 255) / 256, 256, 0, stream>>>(counter, h->cyl_cell_cnt.as<uint32_t>(), rounded, ncodes);
    int rc = run_scan(h, rounded, ncodes, 0, h->cyl_cell_start.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr, nullptr, stream);
    if (rc != TM_OK) return rc;
    uint32_t total = 0, nlong = 0;
    TM_CUDA(h, cudaMemcpyAsync(&total, h->cyl_cell_start.as<uint32_t>() + ncodes, 4, cudaMemcpyDeviceToHost, stream));
    TM_CUDA(h, cudaMemcpyAsync(&nlong, d_nlong, 4, cudaMemcpyDeviceToHost, stream));
    TM_CUDA(h, cudaStreamSynchronize(stream));
    lap("scan + sizes");
    h->n_long = nlong;
    h->index_entries = total;
    h->n_listed = n_regular - nlong;
    const size_t pool = static_cast<size_t>(total) + 4;
    TM_CUDA(h, h->tileA.ensure(sizeof(float4) * pool));
    TM_CUDA(h, h->tileB.ensure(sizeof(float4) * pool));
    TM_CUDA(h, h->tileI.ensure(sizeof(int32_t) * pool));
    TM_CUDA(h, h->tileLB.ensure(sizeof(float) * pool));
    TM_CUDA(h, h->tile_keys.ensure(sizeof(unsigned long long) * pool));
    // padding entries are copied by the bulk loads (never read): give them defined contents
    TM_CUDA(h, cudaMemsetAsync(h->tileA.p, 0, sizeof(float4) * pool, stream));
    TM_CUDA(h, cudaMemsetAsync(h->tileB.p, 0, sizeof(float4) * pool, stream));
    TM_CUDA(h, cudaMemsetAsync(h->tileI.p, 0, sizeof(int32_t) * pool, stream));
    TM_CUDA(h, cudaMemsetAsync(h->tileLB.p, 0, sizeof(float) * pool, stream));
    TM_CUDA(h, cudaMemsetAsync(counter, 0, sizeof(uint32_t) * ncodes, stream));
    cyl_register_kernel<<<blocks, 256, 0, stream>>>(h->recA.as<float4>(), h->recB.as<float4>(), h->boxlo.as<float4>(),
                                                    h->boxhi.as<float4>(), m, g, 1, counter, h->cyl_cell_start.as<uint32_t>(),
                                                    h->tile_keys.as<unsigned long long>(), h->long_list.as<int32_t>(), d_nlong);
    TM_KCHECK(h, stream, "cyl_register_kernel (fill)");
    lap("alloc + register (fill)");
    TM_CUDA(h, cudaMemsetAsync(d_nlong, 0, sizeof(unsigned int), stream));      // reused as the "voxels with a tile" counter
    const size_t sort_smem = sizeof(unsigned long long) * SORT_MAX * SORT_WARPS;
    TM_CUDA(h, cudaFuncSetAttribute(tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sort_smem)));
    const int sort_blocks = static_cast<int>(std::min<uint32_t>((ncodes + SORT_WARPS - 1) / SORT_WARPS, static_cast<uint32_t>(h->sm_count) * 3));
    tile_sort_kernel<<<sort_blocks, SORT_WARPS * 32, sort_smem, stream>>>(
        h->cyl_cell_start.as<uint32_t>(), h->cyl_cell_cnt.as<uint32_t>(), ncodes, h->tile_keys.as<unsigned long long>(),
        h->recA.as<float4>(), h->recB.as<float4>(), h->near, h->tileA.as<float4>(), h->tileB.as<float4>(),
        h->tileI.as<int32_t>(), h->tileLB.as<float>(), h->cyl_cell_near.as<uint32_t>(), d_nlong);
    TM_KCHECK(h, stream, "tile_sort_kernel");
    unsigned int with_tiles = 0;
    TM_CUDA(h, cudaMemcpyAsync(&with_tiles, d_nlong, sizeof(with_tiles), cudaMemcpyDeviceToHost, stream));
    TM_CUDA(h, cudaStreamSynchronize(stream));
    h->voxels_with_tiles = with_tiles;
    lap("tile sort");
    // (tile_keys is build-time scratch, kept for the next table: cudaFree + cudaMalloc cost more than the build's kernels)
    rc = build_bvh(h, stream, static_cast<int>(n_regular), lo, hi);
    if (rc != TM_OK) return rc;
    lap("bvh");
    h->have_grid = true;
    return TM_OK;
}

// ------------------------------------------------------------------------------------------------
// point side: counting sort by voxel id
// ------------------------------------------------------------------------------------------------
constexpr uint32_t NO_CELL = 0xFFFFFFFFu;
constexpr int32_t OUTSIDE_BIT = static_cast<int32_t>(0x80000000u);

// pass 1: voxel occupancy.  The atomics return nothing (RED): no per-point state is kept between the passes, the
// scatter pass recomputes the voxel id from the coordinates (cheaper than 8 bytes of HBM traffic per point each way).
__device__ __forceinline__ uint32_t point_code(const GridDev &g, float x, float y, float z) {
    const float fx = (x - g.ox) * g.inv_h, fy = (y - g.oy) * g.inv_h, fz = (z - g.oz) * g.inv_h;
    // NaN / Inf fail these comparisons and become outliers
    const bool inside = fx >= 0.f && fx < static_cast<float>(g.nx) && fy >= 0.f && fy < static_cast<float>(g.ny) &&
                        fz >= 0.f && fz < static_cast<float>(g.nz);
    return inside ? voxel_code(g, static_cast<int>(fx), static_cast<int>(fy), static_cast<int>(fz)) : NO_CELL;
}

// Both passes aggregate inside the warp first (MATCH.ANY on the voxel id): lanes that hit the same voxel elect a leader
// that issues ONE atomic for the group.  A randomly ordered cloud gains nothing (no two lanes share a voxel, the match
// costs a few instructions next to an L2 atomic); a spatially coherent one — tiled exports, scan lines, re-labelling a
// sorted cloud — sends up to 32x fewer atomics.
__global__ void __launch_bounds__(256) bin_count_kernel(const float *__restrict__ pts, int64_t n, int64_t row_stride, GridDev g, int nsub,
                                                        uint2 *__restrict__ cells, int32_t *__restrict__ pend_idx,
                                                        unsigned long long *__restrict__ pend_keys,
                                                        uint32_t *__restrict__ brute_slots, DevStats *__restrict__ st) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t base = blockIdx.x * static_cast<int64_t>(blockDim.x) + (threadIdx.x & ~31); base < n; base += stride) {
        const int64_t i = base + lane;
        uint32_t code = NO_CELL - 1u - static_cast<uint32_t>(lane);       // idle lanes: distinct ids that match nobody
        bool valid = false;
        if (i < n) {
            const float *p = pts + i * row_stride;
            const float x = p[0], y = p[1], z = p[2];
            const uint32_t c = point_code(g, x, y, z);
            if (c != NO_CELL) {
                code = c;
                valid = true;
            } else {
                // outside the grid: straight to the tree search; non-finite coordinates cannot be bounded at all and take
                // the exhaustive kernel
                const unsigned int s = atomicAdd(&st->pending, 1u);
                pend_idx[s] = static_cast<int32_t>(i) | OUTSIDE_BIT;
                pend_keys[s] = KEY_NONE;
                if (!(fabsf(x) + fabsf(y) + fabsf(z) < 3.0e38f)) brute_slots[atomicAdd(&st->n_brute, 1u)] = s;
            }
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, code);
        // the warp's sub-cell: a function of the row index only, so that the scatter pass finds the same one
        const size_t cell = (static_cast<size_t>(code) * nsub + (static_cast<uint32_t>(base >> 5) & (nsub - 1))) * CELL_PAD;
        if (valid && lane == __ffs(peers) - 1) atomicAdd(&cells[cell].x, static_cast<uint32_t>(__popc(peers)));
    }
}

// pass 2: each point takes the next free slot of its voxel's run.  A cell is {cursor, run start}: one 64-bit atomic per
// group of lanes advances the cursor by the group size and returns both words.
__global__ void __launch_bounds__(256) bin_scatter_kernel(const float *__restrict__ pts, int64_t n, int64_t row_stride, GridDev g, int nsub,
                                                          uint2 *__restrict__ cells, float4 *__restrict__ sorted) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t base = blockIdx.x * static_cast<int64_t>(blockDim.x) + (threadIdx.x & ~31); base < n; base += stride) {
        const int64_t i = base + lane;
        uint32_t code = NO_CELL - 1u - static_cast<uint32_t>(lane);
        bool valid = false;
        float x = 0.f, y = 0.f, z = 0.f;
        if (i < n) {
            const float *p = pts + i * row_stride;
            x = p[0]; y = p[1]; z = p[2];
            const uint32_t c = point_code(g, x, y, z);
            if (c != NO_CELL) { code = c; valid = true; }
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, code);
        const int leader = __ffs(peers) - 1;
        unsigned long long cell = 0;
        if (valid && lane == leader)
            cell = atomicAdd(reinterpret_cast<unsigned long long *>(
                                 cells + (static_cast<size_t>(code) * nsub + (static_cast<uint32_t>(base >> 5) & (nsub - 1))) * CELL_PAD),
                             static_cast<unsigned long long>(__popc(peers)));
        cell = __shfl_sync(0xffffffffu, cell, leader);
        if (valid) {
            const uint32_t pos = static_cast<uint32_t>(cell >> 32) + static_cast<uint32_t>(cell) + static_cast<uint32_t>(__popc(peers & lt));
            sorted[pos] = make_float4(x, y, z, __int_as_float(static_cast<int>(i)));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// evaluate: persistent warps over work items, TMA-staged tiles, cull + dense evaluation, fused record write
// ------------------------------------------------------------------------------------------------
constexpr int EV_WARPS = 8;
constexpr int EV_CHUNK = 64;          // tile entries per stage: 64 * (16 + 16 + 4) B = 2.25 KB
constexpr int Q_CAP = 96;             // < 32 queued pairs before a push, <= 64 pushed per entry

struct __align__(128) WarpStage {
    float4 A[2][EV_CHUNK];
    float4 B[2][EV_CHUNK];
    int32_t I[2][EV_CHUNK];
    float4 P[PTS_PER_ITEM];               // the item's points {x, y, z, bits(original row)}
    unsigned long long best[PTS_PER_ITEM];
    uint32_t q[Q_CAP];                    // (point slot << 16) | entry position in the stage buffers
};

struct EvalArgs {
    const uint4 *items;
    unsigned int *cursor;
    const float4 *sorted;
    const float4 *tileA, *tileB;
    const int32_t *tileI;
    const float4 *recA, *recB;
    const int32_t *ids;
    const int32_t *special, *aligned, *long_list;
    uint32_t n_special, n_aligned, n_long;
    const float *tileLB;
    float atol, eps, slack, near, reach;
    int32_t *win;                 // winning cylinder row of every certified point, at the point's original row
    int32_t *pend_idx;
    unsigned long long *pend_keys;
    DevStats *st;
};

template <bool GUARD, bool NFMA>
__global__ void __launch_bounds__(EV_WARPS * 32, 4) evaluate_kernel(EvalArgs a) {
    extern __shared__ __align__(128) unsigned char ev_smem[];
    WarpStage *stages = reinterpret_cast<WarpStage *>(ev_smem);
    uint64_t *bars = reinterpret_cast<uint64_t *>(ev_smem + sizeof(WarpStage) * EV_WARPS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    WarpStage &ws = stages[warp];
    uint64_t *bar = bars + 2 * warp;
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncwarp();
    const uint32_t n_items = a.st->work_items;
    uint32_t fill = 0, use = 0;          // chunks issued / consumed by this warp

    auto issue_chunk = [&](uint32_t pool_off, uint32_t cnt) {
        if (lane == 0) {
            const int s = fill & 1;
            const uint32_t cnt4 = (cnt + 3u) & ~3u;
            mbar_expect_tx(&bar[s], cnt4 * 36u);
            bulk_g2s(&ws.A[s][0], a.tileA + pool_off, cnt4 * 16u, &bar[s]);
            bulk_g2s(&ws.B[s][0], a.tileB + pool_off, cnt4 * 16u, &bar[s]);
            bulk_g2s(&ws.I[s][0], a.tileI + pool_off, cnt4 * 4u, &bar[s]);
        }
        ++fill;
    };
    auto fetch = [&]() {
        uint32_t v = 0;
        if (lane == 0) v = atomicAdd(a.cursor, 1u);
        return __shfl_sync(0xffffffffu, v, 0);
    };

    uint32_t cur = fetch();
    uint4 it = cur < n_items ? a.items[cur] : make_uint4(0, 0, 0, 0);
    bool cur_ready = false;
    unsigned long long pairs = 0, culls = 0;
    unsigned int nfar = 0;
    while (cur < n_items) {
        const uint32_t nxt = fetch();
        const uint4 itn = nxt < n_items ? a.items[nxt] : make_uint4(0, 0, 0, 0);
        bool nxt_ready = false;
        const uint32_t pool_off = it.x, tcount = it.y, pbeg = it.z, pcnt = it.w & 0xffu, far_cnt = it.w >> 8;
        if (tcount > 0 && !cur_ready) issue_chunk(pool_off, min(static_cast<uint32_t>(EV_CHUNK), tcount));
        const bool v0 = static_cast<uint32_t>(lane) < pcnt, v1 = static_cast<uint32_t>(lane) + 32u < pcnt;
        const float4 P0 = a.sorted[pbeg + min(static_cast<uint32_t>(lane), pcnt - 1)];
        const float4 P1 = a.sorted[pbeg + min(static_cast<uint32_t>(lane) + 32u, pcnt - 1)];
        ws.P[lane] = P0;
        ws.P[lane + 32] = P1;
        ws.best[lane] = KEY_NONE;
        ws.best[lane + 32] = KEY_NONE;
        __syncwarp();
        float thr0 = __int_as_float(0x7f800000), thr1 = thr0;
        uint32_t qn = 0;

        // evaluate up to 32 queued (point, entry) pairs with the reference arithmetic, full lanes
        auto drain = [&]() {
            __syncwarp();
            const uint32_t n = min(qn, 32u);
            if (static_cast<uint32_t>(lane) < n) {
                const uint32_t e = ws.q[qn - n + lane];
                const uint32_t slot = e >> 16, jj = e & 0xffffu;
                const float4 P = ws.P[slot];
                const float4 ca = (&ws.A[0][0])[jj], cb = (&ws.B[0][0])[jj];
                const uint32_t ci = static_cast<uint32_t>((&ws.I[0][0])[jj]);
                const float d = eval_pair<GUARD, NFMA, false>(P.x, P.y, P.z, ca, cb, a.atol, a.eps, nullptr);
                atomicMin(&ws.best[slot], make_key(d, ci));
            }
            qn -= n;
            pairs += n;
            __syncwarp();
            thr0 = thr_of(ws.best[lane], a.slack);
            thr1 = thr_of(ws.best[lane + 32], a.slack);
        };

        if (tcount > 0) {
            const uint32_t nchunks = (tcount + EV_CHUNK - 1) / EV_CHUNK;
            const bool two = pcnt > 32u;
            for (uint32_t ch = 0; ch < nchunks; ++ch) {
                // keep one copy in flight: the next chunk of this tile, else the next item's first chunk
                if (ch + 1 < nchunks) {
                    issue_chunk(pool_off + (ch + 1) * EV_CHUNK, min(static_cast<uint32_t>(EV_CHUNK), tcount - (ch + 1) * EV_CHUNK));
                } else if (nxt < n_items && itn.y > 0) {
                    issue_chunk(itn.x, min(static_cast<uint32_t>(EV_CHUNK), itn.y));
                    nxt_ready = true;
                }
                const int s = use & 1;
                mbar_wait(&bar[s], (use >> 1) & 1);
                ++use;
                const uint32_t cnt = min(static_cast<uint32_t>(EV_CHUNK), tcount - ch * EV_CHUNK);
                if (two) {
                    for (uint32_t j = 0; j < cnt; ++j) {
                        const float4 ca = ws.A[s][j];
                        const float4 cb = ws.B[s][j];
                        const bool p0 = cull_pass(P0.x, P0.y, P0.z, ca, cb, thr0);           // lanes >= 32 carry duplicates:
                        const bool p1 = v1 && cull_pass(P1.x, P1.y, P1.z, ca, cb, thr1);     // all of P0 is valid when two
                        const uint32_t m0 = __ballot_sync(0xffffffffu, p0), m1 = __ballot_sync(0xffffffffu, p1);
                        if (m0 | m1) {
                            const uint32_t e = static_cast<uint32_t>(s * EV_CHUNK) + j;
                            const uint32_t n0 = __popc(m0);
                            if (p0) ws.q[qn + __popc(m0 & lt)] = (static_cast<uint32_t>(lane) << 16) | e;
                            if (p1) ws.q[qn + n0 + __popc(m1 & lt)] = (static_cast<uint32_t>(lane + 32) << 16) | e;
                            qn += n0 + __popc(m1);
                            while (qn >= 32u) drain();
                        }
                    }
                } else {
                    for (uint32_t j = 0; j < cnt; ++j) {
                        const float4 ca = ws.A[s][j];
                        const float4 cb = ws.B[s][j];
                        const bool p0 = v0 && cull_pass(P0.x, P0.y, P0.z, ca, cb, thr0);
                        const uint32_t m0 = __ballot_sync(0xffffffffu, p0);
                        if (m0) {
                            if (p0) ws.q[qn + __popc(m0 & lt)] = (static_cast<uint32_t>(lane) << 16) | (static_cast<uint32_t>(s * EV_CHUNK) + j);
                            qn += __popc(m0);
                            while (qn >= 32u) drain();
                        }
                    }
                }
                while (qn) drain();      // the queue refers to this stage's buffers: empty it before they are re-armed
                __syncwarp();
            }
            culls += static_cast<unsigned long long>(pcnt) * tcount;
        }
        unsigned long long k0 = ws.best[lane], k1 = ws.best[lane + 32];

        // cylinders spanning too many voxels to be listed: cull test against every point
        for (uint32_t e = 0; e < a.n_long; ++e) {
            const uint32_t ci = static_cast<uint32_t>(a.long_list[e]);
            const float4 ca = a.recA[ci], cb = a.recB[ci];
            if (cull_pass(P0.x, P0.y, P0.z, ca, cb, thr_of(k0, a.slack))) {
                const unsigned long long k = make_key(eval_pair<GUARD, NFMA, false>(P0.x, P0.y, P0.z, ca, cb, a.atol, a.eps, nullptr), ci);
                k0 = k < k0 ? k : k0;
            }
            if (cull_pass(P1.x, P1.y, P1.z, ca, cb, thr_of(k1, a.slack))) {
                const unsigned long long k = make_key(eval_pair<GUARD, NFMA, false>(P1.x, P1.y, P1.z, ca, cb, a.atol, a.eps, nullptr), ci);
                k1 = k < k1 ? k : k1;
            }
        }
        // cylinders that cannot be bounded (non-finite / non-unit axis): evaluated for every point
        for (uint32_t e = 0; e < a.n_special; ++e) {
            const uint32_t ci = static_cast<uint32_t>(a.special[e]);
            const float4 ca = a.recA[ci], cb = a.recB[ci];
            const unsigned long long q0 = make_key(eval_pair<GUARD, NFMA, false>(P0.x, P0.y, P0.z, ca, cb, a.atol, a.eps, nullptr), ci);
            const unsigned long long q1 = make_key(eval_pair<GUARD, NFMA, false>(P1.x, P1.y, P1.z, ca, cb, a.atol, a.eps, nullptr), ci);
            k0 = q0 < k0 ? q0 : k0;
            k1 = q1 < k1 ? q1 : k1;
        }
        // variant A: points exactly on the axis line of an axis-parallel cylinder get NaN from it (and NaN wins)
        if (!GUARD) {
            for (uint32_t e = 0; e < a.n_aligned; ++e) {
                const uint32_t ci = static_cast<uint32_t>(a.aligned[e]);
                const float4 ca = a.recA[ci], cb = a.recB[ci];
                if (on_axis_line(P0.x, P0.y, P0.z, ca, cb)) {
                    const unsigned long long k = make_key(eval_pair<GUARD, NFMA, false>(P0.x, P0.y, P0.z, ca, cb, a.atol, a.eps, nullptr), ci);
                    k0 = k < k0 ? k : k0;
                }
                if (on_axis_line(P1.x, P1.y, P1.z, ca, cb)) {
                    const unsigned long long k = make_key(eval_pair<GUARD, NFMA, false>(P1.x, P1.y, P1.z, ca, cb, a.atol, a.eps, nullptr), ci);
                    k1 = k < k1 ? k : k1;
                }
            }
        }
        pairs += static_cast<unsigned long long>(pcnt) * a.n_special;
        culls += static_cast<unsigned long long>(pcnt) * a.n_long;

        // ---- points the near part cannot certify (noise tail): the FAR part of the tile, one point at a time with the
        //      lanes across entries.  Entries are sorted by lb(V, c) <= dist(p, capsule(c)); the first entry whose lb
        //      exceeds the incumbent ends the search (everything behind it, and every cylinder outside the tile, is
        //      farther than the incumbent); an exhausted tile certifies incumbents <= D_max.
        bool far0 = false, far1 = false;
        {
            const bool need0 = v0 && static_cast<uint32_t>(k0 >> 32) != 0u && !(thr_of(k0, a.slack) <= a.near);
            const bool need1 = v1 && static_cast<uint32_t>(k1 >> 32) != 0u && !(thr_of(k1, a.slack) <= a.near);
            const uint32_t nm0 = __ballot_sync(0xffffffffu, need0), nm1 = __ballot_sync(0xffffffffu, need1);
            if (nm0 | nm1) {
                const uint32_t far_off = pool_off + tcount;
#pragma unroll 1
                for (int k = 0; k < 2; ++k) {
                    uint32_t mask = k ? nm1 : nm0;
                    while (mask) {
                        const int src = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const float4 Pk = k ? P1 : P0;
                        const float px = __shfl_sync(0xffffffffu, Pk.x, src), py = __shfl_sync(0xffffffffu, Pk.y, src),
                                    pz = __shfl_sync(0xffffffffu, Pk.z, src);
                        unsigned long long key = __shfl_sync(0xffffffffu, k ? k1 : k0, src);
                        float thr = thr_of(key, a.slack);                  // NaN while there is no incumbent: nothing is skipped
                        bool cut = false;
                        for (uint32_t base = 0; base < far_cnt; base += 32) {
                            const uint32_t j = base + lane;
                            const float lb = j < far_cnt ? a.tileLB[far_off + j] : __int_as_float(0x7f800000);
                            if (__shfl_sync(0xffffffffu, lb, 0) > thr) { cut = true; break; }
                            unsigned long long lk = KEY_NONE;
                            const bool test = j < far_cnt && !(lb > thr);
                            bool hit = false;
                            if (test) {
                                const float4 ca = a.tileA[far_off + j], cb = a.tileB[far_off + j];
                                hit = cull_pass(px, py, pz, ca, cb, thr);
                                if (hit) lk = make_key(eval_pair<GUARD, NFMA, false>(px, py, pz, ca, cb, a.atol, a.eps, nullptr),
                                                       static_cast<uint32_t>(a.tileI[far_off + j]));
                            }
                            culls += __popc(__ballot_sync(0xffffffffu, test));
                            pairs += __popc(__ballot_sync(0xffffffffu, hit));
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                const unsigned long long other = __shfl_xor_sync(0xffffffffu, lk, o);
                                lk = other < lk ? other : lk;
                            }
                            if (lk < key) { key = lk; thr = thr_of(key, a.slack); }
                        }
                        const bool ok = static_cast<uint32_t>(key >> 32) == 0u || cut || thr <= a.reach;
                        if (lane == src) {
                            if (k) { k1 = key; far1 = ok; } else { k0 = key; far0 = ok; }
                        }
                    }
                }
                nfar += __popc(__ballot_sync(0xffffffffu, far0)) + __popc(__ballot_sync(0xffffffffu, far1));
            }
        }

        // ---- certified points: the winning row goes to the point's ORIGINAL row (a 4-byte scatter into an array that
        //      stays L2 resident); the streaming epilogue kernel turns rows into labels + offsets with coalesced
        //      reads and writes.  The rest join the pending list (tree search) with their incumbent ----
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const bool valid = k ? v1 : v0;
            const float4 P = k ? P1 : P0;
            const unsigned long long key = k ? k1 : k0;
            // NaN incumbent (hi word 0) is final: NaN beats everything.  KEY_NONE gives thr = NaN: not certified.
            const bool done = valid && (static_cast<uint32_t>(key >> 32) == 0u || thr_of(key, a.slack) <= a.near || (k ? far1 : far0));
            const bool pend = valid && !done;
            if (done) a.win[__float_as_int(P.w)] = static_cast<int32_t>(key_index(key));
            const uint32_t pm = __ballot_sync(0xffffffffu, pend);
            if (pm) {
                unsigned int base = 0;
                if (lane == 0) base = atomicAdd(&a.st->pending, static_cast<unsigned int>(__popc(pm)));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (pend) {
                    const unsigned int s = base + __popc(pm & lt);
                    a.pend_idx[s] = __float_as_int(P.w);
                    a.pend_keys[s] = key;
                }
            }
        }
        __syncwarp();                    // ws.P / ws.best are rewritten by the next item
        cur = nxt;
        it = itn;
        cur_ready = nxt_ready;
    }
    if (lane == 0 && (pairs | culls)) {
        atomicAdd(&a.st->pairs_grid, pairs);
        atomicAdd(&a.st->cull_tests, culls);
        if (nfar) atomicAdd(&a.st->far_certified, nfar);
    }
}

// ------------------------------------------------------------------------------------------------
// ring search: one CTA per pending point, shells of voxels around its home voxel.  Latency-optimised: it is the
// fast answer when only a handful of points (noise tail) are left; with many uncertified points (clutter) it steps
// aside and the throughput-optimised tree search (tm_bvh.cu) takes all of them.
// ------------------------------------------------------------------------------------------------
constexpr int RING_MAX = 8;
constexpr unsigned int RING_LIMIT = 32768;     // pending points beyond which the tree search is faster (measured crossover ~50k)
constexpr int RING_WARPS = 4;

struct RingArgs {
    const float *pts;
    int64_t row_stride;
    const int32_t *pend_idx;
    unsigned long long *pend_keys;
    uint8_t *pend_done;            // 1 = this slot is final (the tree search skips it)
    const uint32_t *tile_start, *tile_cnt;
    const float4 *tileA, *tileB;
    const int32_t *tileI;
    float atol, eps;
    DevStats *st;
};

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other < v ? other : v;
    }
    return v;
}

// All tiles of voxels V' with dist(p, box(V')) <= rho - D together hold every cylinder within rho of p (walk from
// the cylinder's nearest point towards p by D).  After the Chebyshev shells 0..k every voxel with box distance < k*h
// has been visited, so an incumbent <= D + k*h is certified.  (D = D_max; shell 0, the home tile, was the tile
// kernel's.)  With an incumbent rho the CTA therefore visits, in ONE pass, the voxels of the shells 1..ceil((rho-D)/h)
// that lie within rho - D of the point; without one it grows the search shell by shell until something is found.
// A pass has two block-wide steps so that no thread ever waits on a chain of dependent loads:
//   gather:  the threads enumerate the candidate voxels, test the box distance and append {first entry, count} of the
//            non-empty tiles to a shared list (block-wide scan of the counts);
//   sweep:   the entries of all listed tiles form one flat index space, thread t takes entries t, t + 256, ... (binary
//            search of the list), runs the capsule cull against its incumbent and evaluates the survivors.
// The threads' winners meet in a warp-shuffle min and a shared 64-bit atomicMin.
constexpr int RING_LIST = RING_WARPS * 32 * 4;   // candidate voxels per gather step (4 per thread)
constexpr int RING_LIST_LOG2 = 9;
static_assert((1 << RING_LIST_LOG2) == RING_LIST, "RING_LIST must be a power of two");

template <bool GUARD, bool NFMA>
__global__ void __launch_bounds__(RING_WARPS * 32) ring_kernel(RingArgs a, GridDev g) {
    __shared__ unsigned long long s_key;
    __shared__ uint32_t s_off[RING_LIST], s_beg[RING_LIST + 1];
    __shared__ uint32_t s_warp[RING_WARPS];
    __shared__ uint32_t s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int n_pend = a.st->pending;
    if (n_pend > RING_LIMIT) return;
    unsigned long long pairs = 0, culls = 0;
    for (unsigned int task = blockIdx.x; task < n_pend; task += gridDim.x) {
        const int32_t ri = a.pend_idx[task];
        if (ri < 0) continue;                                  // outside the grid: already on the exhaustive list
        const float *p = a.pts + static_cast<int64_t>(ri) * a.row_stride;
        const float px = p[0], py = p[1], pz = p[2];
        const int hx = static_cast<int>((px - g.ox) * g.inv_h), hy = static_cast<int>((py - g.oy) * g.inv_h),
                  hz = static_cast<int>((pz - g.oz) * g.inv_h);
        unsigned long long key = a.pend_keys[task];
        bool certified = false;
        int kdone = 0;                                         // shells 0..kdone have been searched
        for (;;) {
            const float thr = thr_of(key, g.slack);
            if (static_cast<uint32_t>(key >> 32) == 0u || thr <= g.reach + kdone * g.h) { certified = true; break; }
            if (kdone >= RING_MAX) break;
            const float need_r = thr - g.reach;                // NaN while there is no incumbent: every voxel is needed
            int ktarget = kdone + 1;
            if (need_r == need_r) ktarget = max(ktarget, min(RING_MAX, static_cast<int>(ceilf(need_r * g.inv_h))));
            const int side = 2 * ktarget + 1, total = side * side * side;
            if (tid == 0) s_key = key;
            unsigned long long lk = KEY_NONE;
            float lthr = thr;
            for (int chunk = 0; chunk < total; chunk += RING_LIST) {
                // ---- gather: 4 consecutive candidates per thread
                uint32_t off[4], cnt[4];
                uint32_t mine = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int idx = chunk + tid * 4 + q;
                    off[q] = 0; cnt[q] = 0;
                    if (idx < total) {
                        const int dx = idx % side - ktarget, dy = (idx / side) % side - ktarget, dz = idx / (side * side) - ktarget;
                        const int x = hx + dx, y = hy + dy, z = hz + dz;
                        if (max(max(abs(dx), abs(dy)), abs(dz)) > kdone && x >= 0 && y >= 0 && z >= 0 && x < g.nx && y < g.ny && z < g.nz) {
                            const float lx = g.ox + x * g.h, ly = g.oy + y * g.h, lz = g.oz + z * g.h;
                            const float gx = fmaxf(fmaxf(lx - px, px - (lx + g.h)), 0.f);
                            const float gy = fmaxf(fmaxf(ly - py, py - (ly + g.h)), 0.f);
                            const float gz = fmaxf(fmaxf(lz - pz, pz - (lz + g.h)), 0.f);
                            if (!(sqrtf(gx * gx + gy * gy + gz * gz) > need_r + g.slack)) {
                                const uint32_t code = voxel_code(g, x, y, z);
                                cnt[q] = a.tile_cnt[code];
                                off[q] = a.tile_start[code];
                            }
                        }
                    }
                    mine += cnt[q];
                }
                // block-wide exclusive scan of the per-thread entry counts
                uint32_t inc = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                if (lane == 31) s_warp[warp] = inc;
                __syncthreads();
                uint32_t wbase = 0;
#pragma unroll
                for (int w = 0; w < RING_WARPS; ++w) wbase += w < warp ? s_warp[w] : 0u;
                uint32_t run = wbase + inc - mine;
#pragma unroll
                for (int q = 0; q < 4; ++q) {          // empty tiles get zero-length ranges: harmless for the search
                    s_off[tid * 4 + q] = off[q];
                    s_beg[tid * 4 + q] = run;
                    run += cnt[q];
                }
                if (tid == RING_WARPS * 32 - 1) { s_beg[RING_LIST] = run; s_total = run; }
                __syncthreads();
                // ---- sweep: flat index space over the listed tiles
                const uint32_t n_ent = s_total;
                for (uint32_t e = tid; e < n_ent; e += RING_WARPS * 32) {
                    int lo = 0, hi = RING_LIST;            // last slot with s_beg[slot] <= e
#pragma unroll
                    for (int it = 0; it < RING_LIST_LOG2; ++it) {
                        const int mid = (lo + hi) >> 1;
                        if (s_beg[mid] <= e) lo = mid; else hi = mid;
                    }
                    const uint32_t pos = s_off[lo] + (e - s_beg[lo]);
                    const float4 ca = a.tileA[pos], cb = a.tileB[pos];
                    ++culls;
                    if (cull_pass(px, py, pz, ca, cb, lthr)) {
                        const float d = eval_pair<GUARD, NFMA, false>(px, py, pz, ca, cb, a.atol, a.eps, nullptr);
                        const unsigned long long kk = make_key(d, static_cast<uint32_t>(a.tileI[pos]));
                        lk = kk < lk ? kk : lk;
                        lthr = thr_of(lk < key ? lk : key, g.slack);
                        ++pairs;
                    }
                }
                __syncthreads();                           // the lists are rewritten by the next chunk
            }
            lk = warp_min_u64(lk);
            if (lane == 0 && lk < key) atomicMin(&s_key, lk);
            __syncthreads();
            key = s_key;
            kdone = ktarget;
            __syncthreads();
        }
        if (tid == 0) {
            a.pend_keys[task] = key;                 // certified or not, the incumbent seeds whatever comes next
            if (certified) { a.pend_done[task] = 1; atomicAdd(&a.st->ring_certified, 1u); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
        culls += __shfl_xor_sync(0xffffffffu, culls, o);
    }
    if (lane == 0 && (pairs | culls)) {
        atomicAdd(&a.st->pairs_ring, pairs);
        atomicAdd(&a.st->cull_tests, culls);
    }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
int label_grid(tm_handle *h, const LabelArgs &a) {
    if (a.n == 0) return TM_OK;
    if (a.n > 0x7fffffffLL) return fail(h, TM_ERR_INVALID, "tm_label_points: more than 2^31-1 points per call%s%s");
    cudaStream_t st = a.stream;
    const float want_cell = a.prm.cell_size > 0.f ? a.prm.cell_size : auto_cell_size(h, a.n);
    if (!h->have_grid || h->grid_cell != want_cell) {
        int rc = build_cylinder_index(h, want_cell, st);
        if (rc != TM_OK) return rc;
    }
    h->stats.mode_used = TM_MODE_GRID;
    if (h->n_listed == 0 && h->n_long == 0) return label_brute(h, a);     // only special cylinders: nothing to prune with

    const float slack = slack_for(h->maxabs, h->grid.h);
    const GridDev g = to_dev(h->grid, slack, h->reach, h->near);
    const uint32_t ncodes = h->grid.ncell_codes;
    const size_t n = static_cast<size_t>(a.n);

    // scratch
    const size_t max_occ = std::min<size_t>(n, ncodes);
    const size_t max_items = n / PTS_PER_ITEM + max_occ + 1;
    // dense clouds (hundreds of points per voxel) queue their atomics on the voxels' counters: split each counter into
    // 2 / 4 / 8 sub-cells.  The density is judged by the voxels that have a tile at all (known per table).
    int nsub = 1;
    {
        const double density = static_cast<double>(n) / std::max<uint32_t>(1u, h->voxels_with_tiles);
        if (density > 640.0) nsub = 8; else if (density > 320.0) nsub = 4; else if (density > 160.0) nsub = 2;
        if (const char *env = getenv("TM_SUBCELLS")) { const int v = atoi(env); if (v == 1 || v == 2 || v == 4 || v == 8) nsub = v; }
    }
    TM_CUDA(h, h->cells.ensure(sizeof(uint2) * CELL_PAD * nsub * (static_cast<size_t>(ncodes) + 1)));
    TM_CUDA(h, h->sorted_pts.ensure(sizeof(float4) * n));
    TM_CUDA(h, h->items.ensure(sizeof(uint4) * max_items));
    TM_CUDA(h, h->pend_idx.ensure(sizeof(int32_t) * n));
    TM_CUDA(h, h->brute_slots.ensure(sizeof(uint32_t) * n));
    TM_CUDA(h, h->keys.ensure(sizeof(unsigned long long) * n));
    if (!a.out_index) TM_CUDA(h, h->win.ensure(sizeof(int32_t) * n));
    TM_CUDA(h, h->dstats.ensure(sizeof(DevStats) + 64));
    TM_CUDA(h, h->pend_done.ensure(n));
    TM_CUDA(h, cudaMemsetAsync(h->pend_done.p, 0, n, st));
    DevStats *dst = h->dstats.as<DevStats>();
    unsigned int *cursor = reinterpret_cast<unsigned int *>(h->dstats.as<unsigned char>() + sizeof(DevStats) + 16);
    TM_CUDA(h, cudaMemsetAsync(h->dstats.p, 0, sizeof(DevStats) + 64, st));
    TM_CUDA(h, cudaMemsetAsync(h->cells.p, 0, sizeof(uint2) * CELL_PAD * nsub * ncodes, st));

    const int pt_blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(h->sm_count) * 32));
    bin_count_kernel<<<pt_blocks, 256, 0, st>>>(a.pts, a.n, a.row_stride, g, nsub, h->cells.as<uint2>(),
                                                h->pend_idx.as<int32_t>(), h->keys.as<unsigned long long>(),
                                                h->brute_slots.as<uint32_t>(), dst);
    TM_KCHECK(h, st, "bin_count_kernel");
    mark(h, 1, st);
    int rc = run_scan(h, h->cells.as<uint32_t>(), ncodes, 1, h->cells.as<uint32_t>(), h->cyl_cell_start.as<uint32_t>(),
                      h->cyl_cell_cnt.as<uint32_t>(), h->cyl_cell_near.as<uint32_t>(), h->items.as<uint4>(), dst, st, nsub);
    if (rc != TM_OK) return rc;
    TM_KCHECK(h, st, "scan kernels");
    mark(h, 2, st);
    bin_scatter_kernel<<<pt_blocks, 256, 0, st>>>(a.pts, a.n, a.row_stride, g, nsub, h->cells.as<uint2>(), h->sorted_pts.as<float4>());
    TM_KCHECK(h, st, "bin_scatter_kernel");

    mark(h, 3, st);
    EvalArgs ev;
    ev.items = h->items.as<uint4>();
    ev.cursor = cursor;
    ev.sorted = h->sorted_pts.as<float4>();
    ev.tileA = h->tileA.as<float4>(); ev.tileB = h->tileB.as<float4>(); ev.tileI = h->tileI.as<int32_t>();
    ev.recA = h->recA.as<float4>(); ev.recB = h->recB.as<float4>();
    ev.special = h->special.as<int32_t>();
    ev.aligned = h->aligned.as<int32_t>();
    ev.long_list = h->long_list.as<int32_t>();
    ev.n_special = h->n_special; ev.n_aligned = h->n_aligned; ev.n_long = h->n_long;
    ev.atol = a.prm.perp_atol; ev.eps = a.prm.norm_eps;
    ev.tileLB = h->tileLB.as<float>();
    ev.slack = slack; ev.near = h->near; ev.reach = h->reach;
    int32_t *win = a.out_index ? a.out_index : h->win.as<int32_t>();       // the caller's index array doubles as the scatter target
    ev.win = win;
    ev.pend_idx = h->pend_idx.as<int32_t>();
    ev.pend_keys = h->keys.as<unsigned long long>();
    ev.st = dst;
    const size_t ev_smem = sizeof(WarpStage) * EV_WARPS + sizeof(uint64_t) * 2 * EV_WARPS;
    const bool guard = a.prm.norm_eps > 0.f, nfma = a.prm.norm_fma != 0;
    const int ev_blocks = h->sm_count * 4;
#define TM_EVAL_CASE(G, F)                                                                                         \
    do {                                                                                                           \
        TM_CUDA(h, cudaFuncSetAttribute(evaluate_kernel<G, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                        static_cast<int>(ev_smem)));                                               \
        evaluate_kernel<G, F><<<ev_blocks, EV_WARPS * 32, ev_smem, st>>>(ev);                                      \
    } while (0)
    if (guard) { if (nfma) TM_EVAL_CASE(true, true); else TM_EVAL_CASE(true, false); }
    else       { if (nfma) TM_EVAL_CASE(false, true); else TM_EVAL_CASE(false, false); }
#undef TM_EVAL_CASE
    TM_KCHECK(h, st, "evaluate_kernel");

    // still uncertified at D_max (beyond the far part of their own tile): a handful of points -> ring search, one CTA per
    // point; many (clutter), or outside the grid -> per-point descent of the bounding-volume hierarchy
    mark(h, 4, st);
    RingArgs rg;
    rg.pts = a.pts; rg.row_stride = a.row_stride;
    rg.pend_idx = h->pend_idx.as<int32_t>();
    rg.pend_keys = h->keys.as<unsigned long long>();
    rg.pend_done = h->pend_done.as<uint8_t>();
    rg.tile_start = h->cyl_cell_start.as<uint32_t>(); rg.tile_cnt = h->cyl_cell_cnt.as<uint32_t>();
    rg.tileA = h->tileA.as<float4>(); rg.tileB = h->tileB.as<float4>(); rg.tileI = h->tileI.as<int32_t>();
    rg.atol = a.prm.perp_atol; rg.eps = a.prm.norm_eps;
    rg.st = dst;
    const int rg_blocks = h->sm_count * (2048 / (RING_WARPS * 32));
    if (guard) { if (nfma) ring_kernel<true, true><<<rg_blocks, RING_WARPS * 32, 0, st>>>(rg, g); else ring_kernel<true, false><<<rg_blocks, RING_WARPS * 32, 0, st>>>(rg, g); }
    else       { if (nfma) ring_kernel<false, true><<<rg_blocks, RING_WARPS * 32, 0, st>>>(rg, g); else ring_kernel<false, false><<<rg_blocks, RING_WARPS * 32, 0, st>>>(rg, g); }
    TM_KCHECK(h, st, "ring_kernel");
    rc = search_bvh(h, a, dst);
    if (rc != TM_OK) return rc;

    // exhaustive search for non-finite points, then the winning rows of every pending point
    mark(h, 5, st);
    rc = finish_pending(h, a, dst, win, h->maxabs);
    if (rc != TM_OK) return rc;

    // winner-only epilogue of every row: label + offset, streaming
    mark(h, 7, st);
    return finalize_rows(h, a, win);
}

}  // namespace tmn

// Voxel-grid path: exact nearest-cylinder search with spatial pruning.
//
//   per table (tm_set_cylinders + first use of a cell size):
//     every voxel V gets a TILE: the cylinders whose capsule comes within D_max of box(V), sorted by
//     lb(V, c) = a lower bound of dist(box(V), capsule(c)), packed contiguously ({float4 A, float4 B} | row | lb).
//     The leading `near(V)` entries are those with lb <= D_near.
//   per call:
//     points --count/scan/scatter--> counting sort by voxel id (brick-Morton order) --> contiguous per-voxel runs
//     evaluate:  every lane owns a (voxel, <= 2 points) slot of the sorted cloud and walks the NEAR part of its own
//                voxel's tile — the tiles a warp's 32 slots touch are staged into shared memory with bulk asynchronous
//                copies (TMA) each round, lanes of one voxel read the same words —, computing for every
//                entry a ~30-instruction closed-form distance ESTIMATE in cylinder-local coordinates.  When the best
//                estimate beats the runner-up by more than twice the rounding allowance, no entry is numerically
//                delicate, and the best is within D_near, the winner is decided without a single reference-order
//                evaluation (the epilogue computes the winner's distance and offset in reference order anyway).
//                The other points (~6 %: near-ties, interior points, the noise tail) take the exact kernel: a walk over
//                the tile's entries in ascending lower-bound order, capsule cull against the estimate, survivors
//                queued and evaluated 32 at a time with the reference arithmetic, 64-bit (distance, row) keys —
//                torch.argmin's comparator.  A point whose exact best distance is <= D_near is CERTIFIED: every
//                cylinder that could beat or tie it has lb <= D_near for the point's voxel, i.e. sits in the near
//                part; the noise tail walks on into the FAR part and stops at the first entry whose lb exceeds the
//                incumbent.  The winning row is stored at the point's original row (4 bytes, L2 resident).
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
#include <array>
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
                 const float4 *__restrict__ recB, float near_reach, float4 *__restrict__ tileAB,
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
                tileAB[2 * static_cast<size_t>(off + i)] = recA[c];
                tileAB[2 * static_cast<size_t>(off + i) + 1] = recB[c];
                tileI[off + i] = static_cast<int32_t>(c);
                tileLB[off + i] = lb;
                near += lb <= near_reach ? 1u : 0u;
            }
            __syncwarp();
        } else {
            for (uint32_t i = lane; i < n; i += 32) {
                const uint32_t c = static_cast<uint32_t>(tile_keys[off + i]);
                tileAB[2 * static_cast<size_t>(off + i)] = recA[c];
                tileAB[2 * static_cast<size_t>(off + i) + 1] = recB[c];
                tileI[off + i] = static_cast<int32_t>(c);
                tileLB[off + i] = 0.f;
            }
            near = lane == 0 ? n : 0u;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) near += __shfl_xor_sync(0xffffffffu, near, o);
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

// per voxel code {tile offset, near length, tile length, 0}: what the direct path gathers once per point
__global__ void tile_desc_kernel(const uint32_t *__restrict__ start, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ near,
                                 uint4 *__restrict__ desc, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) desc[i] = make_uint4(start[i], near[i], cnt[i], 0u);
}

// ------------------------------------------------------------------------------------------------
// generic 3-channel exclusive scan over the voxel arrays (points, occupied flags, work items)
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;     // 4096 voxels per block
constexpr int PTS_PER_LANE = 2;           // points of ONE voxel a lane of the tile kernel owns (a "lane slot")
constexpr int CELL_PAD = 4;               // uint2 slots per voxel cell for large clouds: one 32-byte sector each, so that neighbouring voxels'
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
    return Tri{cnt, cnt ? 1u : 0u, (cnt + PTS_PER_LANE - 1) / PTS_PER_LANE};
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
        st->work_items = carry.b;
        st->lane_slots = carry.c;
    }
}

// phase C: final offsets.  mode 0 (tile pool): start[code] only (+ the grand total at start[ncodes]).
// In mode 1 `count` and `start` are the SAME array of {count, start} cells (stride 2): a thread reads the counts of its
// own cells, then rewrites them as {0, start} — the low word becomes the scatter pass's cursor, so that ONE 64-bit
// atomicAdd returns both the run start and the slot inside the run.
// mode 1 (points): start[code] = first sorted point of the voxel; every occupied voxel becomes one item
// (32 bytes: {tile offset, near length, first point, point count | far length, first lane slot, -, -}), in voxel-id order, and every
// group of 32 lane slots learns which item its first slot belongs to (warp_item).
template <int NSUB>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t *count, int stride, uint32_t ncodes,
                                                                  const Tri *__restrict__ block_sums, int mode,
                                                                  uint32_t *start,
                                                                  const uint32_t *__restrict__ tile_start,
                                                                  const uint32_t *__restrict__ tile_cnt,
                                                                  const uint32_t *__restrict__ tile_near,
                                                                  uint4 *__restrict__ items,
                                                                  uint32_t *__restrict__ warp_item) {
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
        uint32_t q_lo = 1, q_hi = 0;                 // groups of 32 lane slots that START inside this voxel's slots
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
                    const uint32_t tnear = tile_near[base + i];
                    items[2 * run.b] = make_uint4(tile_start[base + i], tnear, run.a, cnt[i]);
                    items[2 * run.b + 1] = make_uint4(tile_cnt[base + i] - tnear, run.c, 0u, 0u);
                    const uint32_t nslots = (cnt[i] + PTS_PER_LANE - 1) / PTS_PER_LANE;
                    q_lo = (run.c + 31u) >> 5;
                    q_hi = (run.c + nslots - 1u) >> 5;
                }
            }
        }
        // a few groups -> this thread; a crowded voxel (dense clouds: thousands of points) -> the whole warp
        if (q_lo <= q_hi && q_hi - q_lo < 4u)
            for (uint32_t q = q_lo; q <= q_hi; ++q) warp_item[q] = run.b;
        uint32_t crowded = __ballot_sync(0xffffffffu, q_lo <= q_hi && q_hi - q_lo >= 4u);
        while (crowded) {
            const int src = __ffs(crowded) - 1;
            crowded &= crowded - 1;
            const uint32_t s_lo = __shfl_sync(0xffffffffu, q_lo, src), s_hi = __shfl_sync(0xffffffffu, q_hi, src),
                           s_item = __shfl_sync(0xffffffffu, run.b, src);
            for (uint32_t q = s_lo + lane; q <= s_hi; q += 32) warp_item[q] = s_item;
        }
        run = tri_add(run, tri_of_count(cnt[i]));
    }
    if (mode == 0 && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) start[ncodes] = run.a;
}

static int run_scan(tm_handle *h, const uint32_t *count, uint32_t ncodes, int mode, uint32_t *start, const uint32_t *tile_start,
                    const uint32_t *tile_cnt, const uint32_t *tile_near, uint4 *items, uint32_t *warp_item,
                    DevStats *st, cudaStream_t stream, int nsub = 1, int pad = CELL_PAD) {
    const uint32_t nblocks = (ncodes + SCAN_BLOCK - 1) / SCAN_BLOCK;
    TM_CUDA(h, h->block_sums.ensure(sizeof(Tri) * nblocks));
    Tri *bs = h->block_sums.as<Tri>();
    const int stride = mode == 1 ? 2 * pad : 1;
#define TM_SCAN_CASE(S)                                                                                                        \
    do {                                                                                                                       \
        scan_reduce_kernel<S><<<nblocks, SCAN_THREADS, 0, stream>>>(count, stride, ncodes, bs);                                \
        scan_blocks_kernel<<<1, SCAN_THREADS, 0, stream>>>(bs, nblocks, st);                                                   \
        scan_apply_kernel<S><<<nblocks, SCAN_THREADS, 0, stream>>>(count, stride, ncodes, bs, mode, start, tile_start, tile_cnt, \
                                                                   tile_near, items, warp_item);                               \
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
    return run_scan(h, count, n, 0, start, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
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
    int rc = run_scan(h, rounded, ncodes, 0, h->cyl_cell_start.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
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
    TM_CUDA(h, h->tileAB.ensure(sizeof(float4) * 2 * pool));
    TM_CUDA(h, h->tileI.ensure(sizeof(int32_t) * pool));
    TM_CUDA(h, h->tileLB.ensure(sizeof(float) * pool));
    TM_CUDA(h, h->tile_keys.ensure(sizeof(unsigned long long) * pool));
    // padding entries are copied by the bulk loads (never read): give them defined contents
    TM_CUDA(h, cudaMemsetAsync(h->tileAB.p, 0, sizeof(float4) * 2 * pool, stream));
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
        h->recA.as<float4>(), h->recB.as<float4>(), h->near, h->tileAB.as<float4>(),
        h->tileI.as<int32_t>(), h->tileLB.as<float>(), h->cyl_cell_near.as<uint32_t>(), d_nlong);
    TM_KCHECK(h, stream, "tile_sort_kernel");
    unsigned int with_tiles = 0;
    TM_CUDA(h, cudaMemcpyAsync(&with_tiles, d_nlong, sizeof(with_tiles), cudaMemcpyDeviceToHost, stream));
    TM_CUDA(h, cudaStreamSynchronize(stream));
    h->voxels_with_tiles = with_tiles;
    TM_CUDA(h, h->tile_desc.ensure(sizeof(uint4) * static_cast<size_t>(ncodes)));
    tile_desc_kernel<<<(ncodes + 255) / 256, 256, 0, stream>>>(h->cyl_cell_start.as<uint32_t>(), h->cyl_cell_cnt.as<uint32_t>(),
                                                             h->cyl_cell_near.as<uint32_t>(), h->tile_desc.as<uint4>(), ncodes);
    TM_KCHECK(h, stream, "tile_desc_kernel");
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
constexpr int32_t DONE_BIT = 0x40000000;          // pending slot already final (set by the ring search); rows stay below 2^30

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
__global__ void __launch_bounds__(256) bin_count_kernel(const float *__restrict__ pts, int64_t n, int64_t row_stride, GridDev g, int nsub, int pad,
                                                        uint2 *__restrict__ cells, int32_t *__restrict__ pend_idx,
                                                        unsigned long long *__restrict__ pend_keys,
                                                        uint32_t *__restrict__ brute_slots, int32_t *__restrict__ win,
                                                        DevStats *__restrict__ st) {
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
                win[i] = 0;                                             // provisional: the overlapped epilogue may read it
                if (!(fabsf(x) + fabsf(y) + fabsf(z) < 3.0e38f)) brute_slots[atomicAdd(&st->n_brute, 1u)] = s;
            }
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, code);
        // the warp's sub-cell: a function of the row index only, so that the scatter pass finds the same one
        const size_t cell = (static_cast<size_t>(code) * nsub + (static_cast<uint32_t>(base >> 5) & (nsub - 1))) * pad;
        if (valid && lane == __ffs(peers) - 1) atomicAdd(&cells[cell].x, static_cast<uint32_t>(__popc(peers)));
    }
}

// pass 2: each point takes the next free slot of its voxel's run.  A cell is {cursor, run start}: one 64-bit atomic per
// group of lanes advances the cursor by the group size and returns both words.
__global__ void __launch_bounds__(256) bin_scatter_kernel(const float *__restrict__ pts, int64_t n, int64_t row_stride, GridDev g, int nsub, int pad,
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
                                 cells + (static_cast<size_t>(code) * nsub + (static_cast<uint32_t>(base >> 5) & (nsub - 1))) * pad),
                             static_cast<unsigned long long>(__popc(peers)));
        cell = __shfl_sync(0xffffffffu, cell, leader);
        if (valid) {
            const uint32_t pos = static_cast<uint32_t>(cell >> 32) + static_cast<uint32_t>(cell) + static_cast<uint32_t>(__popc(peers & lt));
            sorted[pos] = make_float4(x, y, z, __int_as_float(static_cast<int>(i)));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tile kernel: lane-owned (voxel, <= 2 points) slots, approximate-distance bounds over the near part of the voxel's
// tile, reference-order evaluations only where the bounds cannot decide
// ------------------------------------------------------------------------------------------------
// The reference distance of a regular cylinder (unit axis, finite record) is, in exact arithmetic,
//     inside the slab (|d| <= atol):  sqrt((rho - r)^2 + d^2)          d   = axial overshoot  t - clamp(t, 0, L)
//     beyond a cap    (|d| >  atol):  sqrt(max(rho - r, 0)^2 + d^2)    rho = distance from the axis line
// and what the reference computes in fp32 differs from it by a few ulp of the largest coordinate in play, which `slack`
// (= S below) bounds with a wide margin.  bound_pair() evaluates that closed form in coordinates relative to the
// cylinder's start (a few ulp of the LOCAL magnitudes, far inside S) and reports
//     D      squared distance estimate (+inf for an unreliable entry)
//     unrel  the estimate cannot be trusted to S: the point is inside the radius where the two forms differ and
//            |d| is within S of atol (for variant A this is every interior point of the slab: the reference's d is
//            rounding noise of the size of its atol = 1e-6), or the point is within S of the axis line (rho -> 0: the
//            reference divides by rho; variant A yields NaN there, which wins the argmin).
// With b = argmin D, m1 = D_b, m2 = min_{j != b} D_j over the reliable entries, no unreliable entry and
// sqrt(m2) - sqrt(m1) > 2 S:
//     ref_b <= sqrt(m1) + S < sqrt(m2) - S <= ref_j   for every j != b,
// so b is the reference's argmin among the tile's near entries with no tie, and it is the argmin over ALL cylinders when
// (sqrt(m1) + S) * 1.00001 + slack <= D_near (everything outside the near part is farther than D_near from every point of
// the voxel).  The winner's distance and offset are computed in reference order by the epilogue kernel, so such a point
// costs no reference-order evaluation here at all.  Every other point (~8 % of a noisy surface cloud: near-ties at
// branch junctions, interior points, the noise tail beyond D_near) takes the exact path: capsule cull against
// thr = sqrt(m1) + S, reference-order evaluation of the survivors, 64-bit (distance, row) keys — the same machinery and
// the same guarantees as before.
constexpr float EST_ALLOWANCE = 0.25f;      // S of the estimate comparison as a fraction of the rounding slack
constexpr int EV_WARPS = 8;
constexpr int Q_CAP = 96;             // < 32 queued pairs before a push, <= 32 pushed per step, drained in 32s
constexpr uint32_t Q_REC = 0x80000000u;    // queue entry refers to a cylinder ROW (recA / recB) instead of a pool position

struct EvalArgs {
    const uint4 *items;           // two uint4 per occupied voxel: {tile offset, near, first point, points} {far, first lane slot, -, -}
    const uint32_t *warp_item;
    unsigned int *cursor;         // next chunk of rounds the tile kernel hands out
    const float4 *sorted;
    const float4 *tileAB;
    const int32_t *tileI;
    const float4 *recA, *recB;
    const int32_t *special, *aligned, *long_list;
    uint32_t n_special, n_aligned, n_long;
    const float *tileLB;
    float atol, eps, slack, near, reach;
    float amb;                    // S of the estimate comparison (<= slack)
    uint4 *undecided;             // points the estimates could not decide: {sorted position, item, bits(upper bound), 0};
    uint32_t undecided_cap;       //   near-certified ones fill the array from the front, the others from the back
    int32_t *win;                 // winning cylinder row of every certified point, at the point's original row
    int32_t *pend_idx;
    unsigned long long *pend_keys;
    DevStats *st;
    // direct path: the records of `undecided` are {row, voxel code, bits(upper bound), 0}
    const float *pts;
    int64_t row_stride;
    const uint4 *tile_desc;       // per voxel code {tile offset, near length, tile length, 0}
    uint32_t far_warp_below, far_wide_below, near_wide_below;     // exact kernel: lanes per walk by list length (see EXACT_*)
};

struct Track {                    // per point: smallest and second smallest squared estimate, entry of the smallest
    float m1, m2;                 // m2 < 0: some entry was unreliable
    uint32_t bj;
};

// WIDE: atol > 2 S, i.e. there is a band |d| < atol - S in which the reference's perp decision is certainly `true`
// (variant B, atol = 1e-3); otherwise (variant A) every |d| <= atol + S is undecided.
template <bool WIDE>
__device__ __forceinline__ void bound_pair(float px, float py, float pz, const float4 A, const float4 B, float band_lo,
                                           float band_hi, float S, float rho2_min, uint32_t j, Track &tr) {
    const float vx = px - A.x, vy = py - A.y, vz = pz - A.z;
    const float t = fmaf(vz, B.z, fmaf(vy, B.y, vx * B.x));
    const float ad = fmaxf(fmaxf(-t, t - A.w), 0.f);         // |t - clamp(t, 0, L)|: one subtraction and one 3-input max (FMNMX3)
    const float rx = fmaf(-t, B.x, vx), ry = fmaf(-t, B.y, vy), rz = fmaf(-t, B.z, vz);     // rejection from the axis LINE
    const float rho2 = fmaf(rz, rz, fmaf(ry, ry, fmaf(rx, rx, 1e-30f)));       // + 1e-30: rsqrt stays finite on the axis line
    const float rho = rho2 * mufu_rsq(rho2);
    const float a = rho - B.w;
    const bool beyond = ad > band_hi;                       // certainly not perp
    const bool undecided = WIDE ? (!beyond && ad >= band_lo) : !beyond;
    // inside the slab the form keeps the sign of a; without the certain band (!WIDE) every slab entry with a < S is
    // unreliable anyway, so max(a, 0) serves both cases there
    const float asel = WIDE ? (beyond ? fmaxf(a, 0.f) : a) : fmaxf(a, 0.f);
    const bool unrel = (undecided && a < S) || rho2 < rho2_min;
    // an unreliable entry gives no upper bound either (variant B divides by max(rho, 1e-8): with rho below the guard the
    // foot point collapses towards the axis and the distance grows to sqrt(d^2 + r^2)): it only forces the exact path
    const float D = unrel ? __int_as_float(0x7f800000) : fmaf(asel, asel, ad * ad);
    tr.m2 = fminf(tr.m2, fmaxf(tr.m1, D));
    tr.bj = D < tr.m1 ? j : tr.bj;
    tr.m1 = fminf(tr.m1, D);
    tr.m2 = unrel ? -1.f : tr.m2;
}

// warp-aggregated append of the lanes with `want` to a list that grows up (dir = +1, from `origin`) or down (dir = -1);
// returns the lane's slot
__device__ __forceinline__ uint32_t warp_append(bool want, unsigned int *counter, uint32_t lane, uint32_t lt) {
    const uint32_t m = __ballot_sync(0xffffffffu, want);
    uint32_t base = 0;
    if (m) {
        if (lane == 0) base = atomicAdd(counter, static_cast<unsigned int>(__popc(m)));
        base = __shfl_sync(0xffffffffu, base, 0);
    }
    return base + __popc(m & lt);
}

constexpr int EV_BLOCKS_PER_SM = PTS_PER_LANE > 2 ? 3 : 4;
constexpr uint32_t EV_CHUNK_ROUNDS = 4;        // consecutive rounds (groups of 32 lane slots) a warp takes per cursor fetch

constexpr uint32_t EV_STAGE_CAP = 128;         // tile entries a warp stages per round: 128 x (32 + 4) B (dynamic shared memory;
                                               // with the 16 KB of append staging 4 CTAs still fit an SM)
constexpr size_t EV_STAGE_BYTES = (sizeof(float4) * 2 + sizeof(int32_t)) * EV_STAGE_CAP;      // 4608 B per warp

struct __align__(16) StageScratch {           // undecided points wait here until 32 of a kind can be written with one atomic
    uint4 front[64];
    uint4 back[64];
};

template <bool WIDE>
__global__ void __launch_bounds__(EV_WARPS * 32, EV_BLOCKS_PER_SM) evaluate_kernel(EvalArgs a) {
    __shared__ StageScratch stage[EV_WARPS];
    extern __shared__ __align__(128) unsigned char tile_stage_raw[];          // per warp: 2 x EV_STAGE_CAP float4 records, then EV_STAGE_CAP rows
    __shared__ __align__(8) uint64_t tile_bar[EV_WARPS];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    StageScratch &sg = stage[threadIdx.x >> 5];
    unsigned char *stage_base = tile_stage_raw + static_cast<size_t>(threadIdx.x >> 5) * EV_STAGE_BYTES;
    float4 *stage_tile = reinterpret_cast<float4 *>(stage_base);
    int32_t *stage_rows = reinterpret_cast<int32_t *>(stage_base + sizeof(float4) * 2 * EV_STAGE_CAP);
    uint64_t *bar = &tile_bar[threadIdx.x >> 5];
    uint32_t bar_phase = 0;
    if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
    const uint32_t n_items = a.st->work_items, total_slots = a.st->lane_slots;
    const uint32_t n_wslots = (total_slots + 31u) >> 5;
    // consecutive rounds per cursor fetch: 4 when every warp still gets several chunks, fewer for small clouds (a warp
    // runs its rounds one after the other: short calls want them spread over all warps)
    const uint32_t n_warps = gridDim.x * EV_WARPS;
    const uint32_t chunk_rounds = n_wslots >= 16u * n_warps ? EV_CHUNK_ROUNDS : (n_wslots >= 4u * n_warps ? 2u : 1u);
    const uint32_t n_chunks = (n_wslots + chunk_rounds - 1u) / chunk_rounds;
    const float S = a.amb;
    const float band_hi = a.atol + S, band_lo = a.atol - S, rho2_min = S * S;
    const bool lists = (a.n_special | a.n_long | a.n_aligned) != 0u;     // warp-uniform: every point takes the exact kernel
    const float INF = __int_as_float(0x7f800000);

    auto fetch = [&]() {
        uint32_t v = 0;
        if (lane == 0) v = atomicAdd(a.cursor, 1u);
        return __shfl_sync(0xffffffffu, v, 0);
    };
    uint32_t nf = 0, nb = 0;                  // staged undecided points (warp-uniform)
    auto flush = [&](uint4 *buf, uint32_t &cnt, unsigned int *counter, bool down) {
        __syncwarp();
        const uint32_t n = min(cnt, 32u);
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(counter, n);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < n) a.undecided[down ? a.undecided_cap - 1u - (base + lane) : base + lane] = buf[lane];
        __syncwarp();
        const uint4 rest = buf[32 + lane];
        __syncwarp();
        if (32u + lane < cnt) buf[lane] = rest;
        cnt -= n;
        __syncwarp();
    };
    auto stage_append = [&](bool want, const uint4 rec, uint4 *buf, uint32_t &cnt, unsigned int *counter, bool down) {
        const uint32_t m = __ballot_sync(0xffffffffu, want);
        if (m) {
            if (want) buf[cnt + __popc(m & lt)] = rec;
            cnt += __popc(m);
            if (cnt >= 32u) flush(buf, cnt, counter, down);
        }
    };

    unsigned int bounds = 0;
    uint32_t chunk = fetch();
    while (chunk < n_chunks) {
        const uint32_t next_chunk = fetch();
        const uint32_t r_end = min((chunk + 1u) * chunk_rounds, n_wslots);
        for (uint32_t cur = chunk * chunk_rounds; cur < r_end; ++cur) {
            // ---- which voxel run does this lane's slot belong to?  The warp's 32 slots span at most 32 consecutive items.
            const uint32_t it0 = a.warp_item[cur];
            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0xFFFFFFFFu, 0, 0);       // lane l holds the record of item it0 + l
            if (it0 + lane < n_items) { r0 = a.items[2 * (it0 + lane)]; r1 = a.items[2 * (it0 + lane) + 1]; }
            const uint32_t s = (cur << 5) + lane;
            const bool valid = s < total_slots;
            uint32_t lo = 0, hi = 32;
#pragma unroll
            for (int step = 0; step < 5; ++step) {
                const uint32_t mid = (lo + hi) >> 1;
                const uint32_t v = __shfl_sync(0xffffffffu, r1.y, mid);
                if (v <= s) lo = mid; else hi = mid;
            }
            // the lane's own item: its record sits in lane `lo` (no second trip to memory)
            const uint32_t item = it0 + lo;
            const uint32_t tile_off = __shfl_sync(0xffffffffu, r0.x, lo);
            const uint32_t near_raw = __shfl_sync(0xffffffffu, r0.y, lo), first_pt = __shfl_sync(0xffffffffu, r0.z, lo),
                           n_pts = __shfl_sync(0xffffffffu, r0.w, lo), far_raw = __shfl_sync(0xffffffffu, r1.x, lo),
                           slot_start = __shfl_sync(0xffffffffu, r1.y, lo);
            const uint32_t near_cnt = valid ? near_raw : 0u;
            const uint32_t k = valid ? s - slot_start : 0u;
            const uint32_t p0 = first_pt + PTS_PER_LANE * k;
            bool pv[PTS_PER_LANE];
#pragma unroll
            for (int q = 0; q < PTS_PER_LANE; ++q) pv[q] = valid && PTS_PER_LANE * k + q < n_pts;
            const float4 *tile = a.tileAB + 2 * static_cast<size_t>(tile_off);
            // ---- the near parts of the tiles this round touches (the warp's slots span n_dist consecutive items), staged
            //      into the warp's shared-memory buffer with bulk asynchronous copies (TMA: records and cylinder rows of every
            //      tile, issued by the lane that holds the item's record) completing on the warp's mbarrier: one request per
            //      tile instead of one L2 round trip per entry in the loop below.  Tiles start on multiples of 4 entries in
            //      the pool and in the buffer (16-byte alignment of the row indices).  Rounds whose tiles do not fit read them
            //      from global memory.
            const uint32_t n_dist = __reduce_max_sync(0xffffffffu, valid ? lo + 1u : 0u);
            const uint32_t st_off = r0.x;
            const uint32_t st_near = lane < n_dist ? (r0.y + 3u) & ~3u : 0u;
            uint32_t st_pos = st_near;                                  // inclusive scan of the near lengths
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, st_pos, o); if (lane >= static_cast<uint32_t>(o)) st_pos += v; }
            const uint32_t st_total = __shfl_sync(0xffffffffu, st_pos, 31);
            st_pos -= st_near;
            const bool staged = st_total > 0u && st_total <= EV_STAGE_CAP;
            if (staged) {
                fence_proxy_async();                                    // the previous round's reads precede these writes
                if (lane == 0) mbar_expect_tx(bar, st_total * 36u);
                __syncwarp();
                if (st_near) {
                    bulk_g2s(stage_tile + 2 * st_pos, a.tileAB + 2 * static_cast<size_t>(st_off), st_near * 32u, bar);
                    bulk_g2s(stage_rows + st_pos, a.tileI + st_off, st_near * 4u, bar);
                }
            }
            const uint32_t my_pos = __shfl_sync(0xffffffffu, st_pos, lo);
            float4 P[PTS_PER_LANE];
            P[0] = pv[0] ? a.sorted[p0] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 1; q < PTS_PER_LANE; ++q) P[q] = pv[q] ? a.sorted[p0 + q] : P[0];
            // ---- estimates over the near part of the lane's own tile, entries double-buffered in registers
            Track tr[PTS_PER_LANE];
#pragma unroll
            for (int q = 0; q < PTS_PER_LANE; ++q) tr[q] = Track{INF, INF, 0u};
            if (staged) {
                mbar_wait(bar, bar_phase);
                bar_phase ^= 1u;
                const float4 *mine_tile = stage_tile + 2 * my_pos;         // lanes of one voxel read the same words: broadcast
                // every lane runs over its own tile's length (lanes of one voxel stay together, the others leave the loop
                // early): no per-entry reconvergence bookkeeping
#pragma unroll 2
                for (uint32_t j = 0; j < near_cnt; ++j) {
                    const float4 A = mine_tile[2 * j], B = mine_tile[2 * j + 1];
#pragma unroll
                    for (int q = 0; q < PTS_PER_LANE; ++q) bound_pair<WIDE>(P[q].x, P[q].y, P[q].z, A, B, band_lo, band_hi, S, rho2_min, j, tr[q]);
                }
                __syncwarp();                                               // everyone is done with the buffer
            } else {
                float4 A0 = make_float4(0.f, 0.f, 0.f, 0.f), B0 = A0, A1 = A0, B1 = A0;
                if (0u < near_cnt) { A0 = tile[0]; B0 = tile[1]; }
                for (uint32_t j = 0; j < near_cnt; j += 2) {
                    if (j + 1u < near_cnt) { A1 = tile[2 * (j + 1u)]; B1 = tile[2 * (j + 1u) + 1]; }
#pragma unroll
                    for (int q = 0; q < PTS_PER_LANE; ++q) bound_pair<WIDE>(P[q].x, P[q].y, P[q].z, A0, B0, band_lo, band_hi, S, rho2_min, j, tr[q]);
                    if (j + 2u < near_cnt) { A0 = tile[2 * (j + 2u)]; B0 = tile[2 * (j + 2u) + 1]; }
                    if (j + 1u < near_cnt) {
#pragma unroll
                        for (int q = 0; q < PTS_PER_LANE; ++q) bound_pair<WIDE>(P[q].x, P[q].y, P[q].z, A1, B1, band_lo, band_hi, S, rho2_min, j + 1u, tr[q]);
                    }
                }
            }
            uint32_t npts = 0;
#pragma unroll
            for (int q = 0; q < PTS_PER_LANE; ++q) npts += pv[q] ? 1u : 0u;
            bounds += near_cnt * npts;

            // ---- decided by the estimates alone?  (d1 = NaN without a reliable entry: every comparison below fails)
            float up[PTS_PER_LANE];
            bool cert[PTS_PER_LANE], todo[PTS_PER_LANE];
            bool any_todo = false;
#pragma unroll
            for (int q = 0; q < PTS_PER_LANE; ++q) {
                const float d1 = tr[q].m1 * mufu_rsq(fmaxf(tr[q].m1, 1e-30f));
                up[q] = fmaf(d1 + S, 1.00001f, a.slack);                     // >= thr of the exact winner
                cert[q] = up[q] <= a.near;
                const float w = d1 + 2.f * S;
                const bool fast = pv[q] && !lists && cert[q] && tr[q].m2 > w * w;      // m2 = -1 (unreliable entry) fails
                if (fast) a.win[__float_as_int(P[q].w)] = staged ? stage_rows[my_pos + tr[q].bj] : a.tileI[tile_off + tr[q].bj];
                todo[q] = pv[q] && !fast;
                any_todo = any_todo || todo[q];
            }

            if (__any_sync(0xffffffffu, any_todo)) {
                // voxels without any tile entry (clutter far from every cylinder) and no list to run: straight to the pending list
                const bool empty = !lists && valid && near_cnt == 0u && far_raw == 0u;
#pragma unroll
                for (int q = 0; q < PTS_PER_LANE; ++q) {
                    const uint4 rec = make_uint4(p0 + q, item, __float_as_uint(up[q]), 0u);
                    stage_append(todo[q] && !empty && cert[q], rec, sg.front, nf, &a.st->undecided_near, false);
                    stage_append(todo[q] && !empty && !cert[q], rec, sg.back, nb, &a.st->undecided_far, true);
                    const bool pend = todo[q] && empty;
                    const uint32_t sp = warp_append(pend, &a.st->pending, lane, lt);
                    if (pend) {
                        a.pend_idx[sp] = __float_as_int(P[q].w);
                        a.pend_keys[sp] = KEY_NONE;
                        a.win[__float_as_int(P[q].w)] = 0;              // provisional: the overlapped epilogue may read it
                    }
                }
            }
        }
        chunk = next_chunk;
    }
    while (nf) flush(sg.front, nf, &a.st->undecided_near, false);
    while (nb) flush(sg.back, nb, &a.st->undecided_far, true);
    unsigned long long all_bounds = bounds;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) all_bounds += __shfl_xor_sync(0xffffffffu, all_bounds, o);
    if (lane == 0 && all_bounds) atomicAdd(&a.st->bound_tests, all_bounds);
}

// ------------------------------------------------------------------------------------------------
// exact kernel: the points the estimates left undecided, 32 per warp, each lane walking its own point's tile in
// ascending lower-bound order: capsule cull against the point's upper bound, survivors queued per warp and evaluated 32
// at a time with the reference arithmetic (full lanes), 64-bit (distance, row) keys in shared memory.
// Near-certified points (front of the list) stop at the end of the near part; the others (back of the list: the noise
// tail) walk on through the FAR part and stop at the first entry whose lower bound exceeds their incumbent, so the two
// kinds never share a warp and the walks of a warp have similar lengths.
// ------------------------------------------------------------------------------------------------
struct __align__(16) ExactScratch {
    float4 P[32];
    unsigned long long best[32];
    uint2 q[Q_CAP];                           // {pool position (or row | Q_REC), lane of the point}
};

template <bool GUARD, bool NFMA>
__device__ __forceinline__ void exact_drain(const EvalArgs &a, ExactScratch &ws, uint32_t lane, uint32_t &qn, unsigned int &pairs) {
    __syncwarp();
    const uint32_t n = min(qn, 32u);
    if (lane < n) {
        const uint2 e = ws.q[qn - n + lane];
        const float4 P = ws.P[e.y];
        float4 ca, cb;
        uint32_t ci;
        if (e.x & Q_REC) { ci = e.x & ~Q_REC; ca = a.recA[ci]; cb = a.recB[ci]; }
        else { ca = a.tileAB[2 * static_cast<size_t>(e.x)]; cb = a.tileAB[2 * static_cast<size_t>(e.x) + 1]; ci = static_cast<uint32_t>(a.tileI[e.x]); }
        const float d = eval_pair<GUARD, NFMA, false>(P.x, P.y, P.z, ca, cb, a.atol, a.eps, nullptr);
        atomicMin(&ws.best[e.y], make_key(d, ci));
    }
    qn -= n;
    pairs += n;
    __syncwarp();
}

__device__ __forceinline__ void exact_push(ExactScratch &ws, uint32_t lt, bool hit, uint32_t pos, uint32_t slot, uint32_t &qn) {
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    if (m) {
        if (hit) ws.q[qn + __popc(m & lt)] = make_uint2(pos, slot);
        qn += __popc(m);
    }
}

// One task = 32 / GS undecided points, GS lanes per point.  GS = 1: every lane walks its own point's tile (the short walks
// of the near-certified points); GS = 8: eight lanes step through a point's tile together (the noise tail walks the whole
// far part, often hundreds of entries: eight times fewer dependent steps per point).
template <bool GUARD, bool NFMA, bool DIRECT, int GS>
__device__ __forceinline__ void exact_task(const EvalArgs &a, ExactScratch &ws, uint32_t lane, uint32_t lt, uint32_t task,
                                           bool is_front, uint32_t count, uint32_t &qn, unsigned int &pairs,
                                           unsigned int &culls, unsigned int &nfar) {
    constexpr uint32_t PW = 32 / GS;
    const uint32_t g = lane / GS, sub = lane % GS;
    const uint32_t leader = lane - sub;
    const float INF = __int_as_float(0x7f800000);
    const uint32_t idx = task * PW + g;
    const bool valid = idx < count;
    const uint4 rec = valid ? a.undecided[is_front ? idx : a.undecided_cap - 1u - idx] : make_uint4(0, 0, 0, 0);
    float4 P = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t off = 0, total = 0;
    if (valid) {
        if (DIRECT) {
            const float *p = a.pts + static_cast<int64_t>(rec.x) * a.row_stride;
            P = make_float4(p[0], p[1], p[2], __uint_as_float(rec.x));
            const uint4 d = a.tile_desc[rec.y];
            off = d.x;
            total = is_front ? d.y : d.z;
        } else {
            P = a.sorted[rec.x];
            const uint4 it = a.items[2 * rec.y];
            off = it.x;
            total = is_front ? it.y : it.y + a.items[2 * rec.y + 1].x;
        }
    }
    float thr = __uint_as_float(rec.z);            // NaN without a reliable estimate: nothing is culled
    // the walk below reads the tile entry by entry: ask for the first lines now, the rest as the walk advances
    constexpr uint32_t AHEAD = GS > 8 ? 4u * GS : 32u;          // entries requested ahead of the walk (four steps of a wide group)
    if (total) {
        if (sub == 0) prefetch_l1(a.tileLB + off);
        for (uint32_t e = 4 * sub; e < min(total, AHEAD); e += 4 * GS) prefetch_l1(a.tileAB + 2 * static_cast<size_t>(off + e));
    }
    if (sub == 0) { ws.P[g] = P; ws.best[g] = KEY_NONE; }
    __syncwarp();
    bool cut = false;                              // uniform within the group
    uint32_t live = __ballot_sync(0xffffffffu, total > 0u);
    for (uint32_t base = 0; live; base += GS) {
        const uint32_t j = base + sub;
        bool hit = false;
        float lb = INF;
        if (!cut && j < total) {
            if ((j & 3u) == 0u && j + AHEAD < total) prefetch_l1(a.tileAB + 2 * static_cast<size_t>(off + j + AHEAD));
            if ((j & 31u) == 0u && j + 2u * AHEAD < total) prefetch_l1(a.tileLB + off + j + 2u * AHEAD);
            lb = a.tileLB[off + j];
        }
        // the group's first entry of this step beyond the incumbent: everything from here on is farther
        const float lb0 = GS == 1 ? lb : __shfl_sync(0xffffffffu, lb, leader);
        if (!cut && base < total && lb0 > thr) cut = true;
        if (!cut && j < total && !(lb > thr)) {
            const float4 ca = a.tileAB[2 * static_cast<size_t>(off + j)], cb = a.tileAB[2 * static_cast<size_t>(off + j) + 1];
            hit = cull_pass(P.x, P.y, P.z, ca, cb, thr);
            ++culls;
        }
        exact_push(ws, lt, hit, off + j, g, qn);
        if (qn >= 32u) {
            exact_drain<GUARD, NFMA>(a, ws, lane, qn, pairs);
            thr = fminf(thr, thr_of(ws.best[g], a.slack));     // fminf ignores a NaN operand
        }
        live = __ballot_sync(0xffffffffu, !cut && base + GS < total);
    }
    // cylinders that cannot be bounded (non-finite / non-unit axis): evaluated for every point; cylinders spanning too
    // many voxels to be listed: cull test; variant A: points exactly on the axis line of an axis-parallel cylinder get
    // NaN from it (and NaN wins)
    for (uint32_t e0 = 0; e0 < a.n_special; e0 += GS) {
        const uint32_t e = e0 + sub;
        exact_push(ws, lt, valid && e < a.n_special, e < a.n_special ? (static_cast<uint32_t>(a.special[e]) | Q_REC) : 0u, g, qn);
        while (qn >= 32u) exact_drain<GUARD, NFMA>(a, ws, lane, qn, pairs);
    }
    if (a.n_long) {
        while (qn) exact_drain<GUARD, NFMA>(a, ws, lane, qn, pairs);
        thr = fminf(thr, thr_of(ws.best[g], a.slack));
        for (uint32_t e0 = 0; e0 < a.n_long; e0 += GS) {
            const uint32_t e = e0 + sub;
            bool hit = false;
            uint32_t ci = 0;
            if (valid && e < a.n_long) {
                ci = static_cast<uint32_t>(a.long_list[e]);
                hit = cull_pass(P.x, P.y, P.z, a.recA[ci], a.recB[ci], thr);
                ++culls;
            }
            exact_push(ws, lt, hit, ci | Q_REC, g, qn);
            while (qn >= 32u) exact_drain<GUARD, NFMA>(a, ws, lane, qn, pairs);
        }
    }
    if (!GUARD) {
        for (uint32_t e0 = 0; e0 < a.n_aligned; e0 += GS) {
            const uint32_t e = e0 + sub;
            bool hit = false;
            uint32_t ci = 0;
            if (valid && e < a.n_aligned) {
                ci = static_cast<uint32_t>(a.aligned[e]);
                hit = on_axis_line(P.x, P.y, P.z, a.recA[ci], a.recB[ci]);
            }
            exact_push(ws, lt, hit, ci | Q_REC, g, qn);
            while (qn >= 32u) exact_drain<GUARD, NFMA>(a, ws, lane, qn, pairs);
        }
    }
    while (qn) exact_drain<GUARD, NFMA>(a, ws, lane, qn, pairs);
    const unsigned long long key = ws.best[g];
    thr = fminf(thr, thr_of(key, a.slack));
    // NaN incumbent (hi word 0) is final: NaN beats everything.  Near-certified: the exact distance is inside D_near.
    // Otherwise: an entry beyond the incumbent ended the walk, or the whole tile was searched and the incumbent is
    // inside D_max — everything else is farther.  KEY_NONE is never certified.
    const bool mine = valid && sub == 0;
    const bool nan_key = static_cast<uint32_t>(key >> 32) == 0u;
    const bool near_ok = thr_of(key, a.slack) <= a.near;
    const bool far_ok = !is_front && key != KEY_NONE && (cut || thr <= a.reach);
    const bool done = mine && (nan_key || near_ok || far_ok);
    if (done) a.win[__float_as_int(P.w)] = static_cast<int32_t>(key_index(key));
    nfar += __popc(__ballot_sync(0xffffffffu, done && !nan_key && !near_ok));
    const bool pend = mine && !done;
    const uint32_t sp = warp_append(pend, &a.st->pending, lane, lt);
    if (pend) {
        a.pend_idx[sp] = __float_as_int(P.w);
        a.pend_keys[sp] = key;
        a.win[__float_as_int(P.w)] = key == KEY_NONE ? 0 : static_cast<int32_t>(key_index(key));     // provisional
    }
    __syncwarp();                    // ws.P / ws.best are rewritten by the next task
}

// Lanes per point of a walk: a walk is a chain of dependent steps, so a small call (every warp gets a task or two anyway)
// spends more lanes on each point and finishes in fewer steps; a large one is throughput work, one lane per point.
constexpr int EXACT_FAR_LANES = 8;
constexpr uint32_t EXACT_FAR_WIDE_BELOW = 100000;       // far walks (hundreds of entries): 8 lanes per point below this many of them,
constexpr uint32_t EXACT_FAR_WARP_BELOW = 4096;         //   a whole warp per point below this
constexpr uint32_t EXACT_NEAR_WIDE_BELOW = 131072;      // near walks (tens of entries): 8 lanes per point below this many

template <bool GUARD, bool NFMA, bool DIRECT>
__global__ void __launch_bounds__(EV_WARPS * 32, 4) exact_kernel(EvalArgs a) {
    __shared__ ExactScratch scratch[EV_WARPS];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    ExactScratch &ws = scratch[warp];
    const uint32_t n_front = a.st->undecided_near, n_back = a.st->undecided_far;
    // grid-uniform choices
    const uint32_t far_lanes = n_back < a.far_warp_below ? 32u : (n_back < a.far_wide_below ? static_cast<uint32_t>(EXACT_FAR_LANES) : 1u);
    const uint32_t near_lanes = n_front < a.near_wide_below ? static_cast<uint32_t>(EXACT_FAR_LANES) : 1u;
    const uint32_t PB = 32u / far_lanes, PF = 32u / near_lanes;
    const uint32_t w_front = (n_front + PF - 1u) / PF, w_all = w_front + (n_back + PB - 1u) / PB;
    unsigned int pairs = 0, culls = 0, nfar = 0;
    uint32_t qn = 0;
    // the long walks first: they are the critical path of a small call
    for (uint32_t w = blockIdx.x * EV_WARPS + warp; w < w_all; w += gridDim.x * EV_WARPS) {
        if (w < w_all - w_front) {
            if (far_lanes == 32u) exact_task<GUARD, NFMA, DIRECT, 32>(a, ws, lane, lt, w, false, n_back, qn, pairs, culls, nfar);
            else if (far_lanes == 1u) exact_task<GUARD, NFMA, DIRECT, 1>(a, ws, lane, lt, w, false, n_back, qn, pairs, culls, nfar);
            else exact_task<GUARD, NFMA, DIRECT, EXACT_FAR_LANES>(a, ws, lane, lt, w, false, n_back, qn, pairs, culls, nfar);
        } else {
            if (near_lanes == 1u) exact_task<GUARD, NFMA, DIRECT, 1>(a, ws, lane, lt, w - (w_all - w_front), true, n_front, qn, pairs, culls, nfar);
            else exact_task<GUARD, NFMA, DIRECT, EXACT_FAR_LANES>(a, ws, lane, lt, w - (w_all - w_front), true, n_front, qn, pairs, culls, nfar);
        }
    }
    unsigned long long all_culls = culls;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) all_culls += __shfl_xor_sync(0xffffffffu, all_culls, o);
    if (lane == 0 && (pairs | all_culls)) {
        atomicAdd(&a.st->pairs_grid, static_cast<unsigned long long>(pairs));
        atomicAdd(&a.st->cull_tests, all_culls);
        if (nfar) atomicAdd(&a.st->far_certified, nfar);
    }
}

// ------------------------------------------------------------------------------------------------
// direct path: no sort at all.  One thread per point in INPUT order: voxel id -> the voxel's tile descriptor (one 16-byte
// gather) -> estimates over the near part of the tile (32-byte gathers that hit L2; lanes of a warp sit in different
// voxels unless the cloud is spatially coherent) -> when the estimates decide, the winner is evaluated in reference
// order on the spot and label + offset are written in input order (coalesced): no counting sort, no scan over the voxels,
// no separate epilogue, no per-call cost that does not scale with the number of points.  Undecided points go to the
// exact kernel, stragglers to the ring / tree search as before, and a short epilogue finishes exactly those rows.
// To keep the lanes of a warp in step, the 256 points of a block are first ordered by the length of their tile's near
// part (counting sort in shared memory): a warp then loops over similar lengths instead of the longest of 32 random ones.
// ------------------------------------------------------------------------------------------------
constexpr int DIRECT_THREADS = 256;
constexpr uint32_t DIRECT_BINS = 64;          // near lengths 0..62 get their own bin, longer ones share the last
constexpr uint32_t DIRECT_AHEAD = 16;         // tile entries requested ahead of the estimate loop (4 lines of 128 bytes)

struct DirectArgs {
    const float *pts;
    int64_t n, row_stride;
    const uint4 *tile_desc;       // per voxel code {tile offset, near length, tile length, 0}
    EvalArgs ev;
    const float4 *recAB;
    const int32_t *ids;
    int move_to_mantle;
    int32_t *out_index, *out_id;
    float *out_dist, *out_offset, *out_radius;
    float4 *out_packed;
    uint32_t *late_rows;          // rows that did not get their outputs here (finished by finalize_list_kernel)
    uint32_t *brute_slots;
};

template <bool GUARD, bool NFMA, bool WIDE>
__global__ void __launch_bounds__(DIRECT_THREADS, 6) direct_kernel(DirectArgs a, GridDev g) {
    __shared__ uint32_t s_hist[DIRECT_BINS];
    __shared__ uint32_t s_perm[DIRECT_THREADS];
    __shared__ float4 s_pt[DIRECT_THREADS];       // {x, y, z, bits(near length)}
    __shared__ uint2 s_tile[DIRECT_THREADS];      // {tile offset | code}
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const EvalArgs &e = a.ev;
    const float S = e.amb;
    const float band_hi = e.atol + S, band_lo = e.atol - S, rho2_min = S * S;
    const bool lists = (e.n_special | e.n_long | e.n_aligned) != 0u;
    const float INF = __int_as_float(0x7f800000);
    unsigned int bounds = 0, evals = 0;
    for (int64_t base = static_cast<int64_t>(blockIdx.x) * DIRECT_THREADS; base < a.n; base += static_cast<int64_t>(gridDim.x) * DIRECT_THREADS) {
        // ---- load, bin, order by near length
        if (tid < DIRECT_BINS) s_hist[tid] = 0;
        __syncthreads();
        const int64_t i = base + tid;
        uint32_t code = NO_CELL, bin = DIRECT_BINS - 1u, rank = 0;
        uint4 desc = make_uint4(0, 0, 0, 0);
        float x = 0.f, y = 0.f, z = 0.f;
        const bool have = i < a.n;
        if (have) {
            const float *p = a.pts + i * a.row_stride;
            x = p[0]; y = p[1]; z = p[2];
            code = point_code(g, x, y, z);
            if (code != NO_CELL) desc = a.tile_desc[code];
            bin = min(desc.y, DIRECT_BINS - 1u);
            // the estimate loop below reads the tile entry by entry, one dependent trip to memory each (a cold table: HBM
            // latency x the longest near part of the block); ask for the first lines now, while the block is being ordered
            // (tiles start on 128-byte lines of 4 entries), the rest as the loop advances
            if (desc.y) {
                prefetch_l1(e.tileI + desc.x);
                for (uint32_t k = 0; k < min(desc.y, DIRECT_AHEAD); k += 4) prefetch_l1(e.tileAB + 2 * static_cast<size_t>(desc.x + k));
            }
        }
        rank = atomicAdd(&s_hist[bin], 1u);
        __syncthreads();
        if (tid < 32) {                                   // exclusive scan of the 64 bins by one warp
            const uint32_t c0 = s_hist[2 * tid], c1 = s_hist[2 * tid + 1];
            uint32_t inc = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= static_cast<uint32_t>(o)) inc += v; }
            s_hist[2 * tid] = inc - c0 - c1;
            s_hist[2 * tid + 1] = inc - c1;
        }
        __syncthreads();
        const uint32_t pos = s_hist[bin] + rank;
        s_perm[pos] = tid;
        s_pt[pos] = make_float4(x, y, z, __uint_as_float(desc.y));
        s_tile[pos] = make_uint2(desc.x, code);
        __syncthreads();
        // ---- this thread now works on the point at position tid of the ordered block
        const uint32_t src = s_perm[tid];
        const int64_t row = base + src;
        const bool valid = row < a.n;
        const float4 P = s_pt[tid];
        const uint2 tl = s_tile[tid];
        const uint32_t near_cnt = valid ? __float_as_uint(P.w) : 0u, tile_off = tl.x, mycode = tl.y;
        const bool inside = valid && mycode != NO_CELL;

        Track t0{INF, INF, 0u};
        const float4 *tile = e.tileAB + 2 * static_cast<size_t>(tile_off);
        float4 A0 = make_float4(0.f, 0.f, 0.f, 0.f), B0 = A0, A1 = A0, B1 = A0;
        if (0u < near_cnt) { A0 = tile[0]; B0 = tile[1]; }
        for (uint32_t j = 0; j < near_cnt; j += 2) {               // per-lane trip count: the lanes meet again after the loop
            if ((j & 2u) == 0u && j + DIRECT_AHEAD < near_cnt) prefetch_l1(tile + 2 * (j + DIRECT_AHEAD));
            if (j + 1u < near_cnt) { A1 = tile[2 * (j + 1u)]; B1 = tile[2 * (j + 1u) + 1]; }
            bound_pair<WIDE>(P.x, P.y, P.z, A0, B0, band_lo, band_hi, S, rho2_min, j, t0);
            if (j + 2u < near_cnt) { A0 = tile[2 * (j + 2u)]; B0 = tile[2 * (j + 2u) + 1]; }
            if (j + 1u < near_cnt) bound_pair<WIDE>(P.x, P.y, P.z, A1, B1, band_lo, band_hi, S, rho2_min, j + 1u, t0);
        }
        bounds += near_cnt;

        const float d1 = t0.m1 * mufu_rsq(fmaxf(t0.m1, 1e-30f));
        const float up = fmaf(d1 + S, 1.00001f, e.slack);
        const bool cert = up <= e.near;
        const float w = d1 + 2.f * S;
        const bool fast = inside && !lists && cert && t0.m2 > w * w;
        if (fast) {
            // the winner, in reference order, and its outputs at the point's own row
            const uint32_t pos_w = tile_off + t0.bj;
            const uint32_t ci = static_cast<uint32_t>(e.tileI[pos_w]);
            const float4 ca = tile[2 * t0.bj], cb = tile[2 * t0.bj + 1];
            PairGeom gm;
            eval_pair<GUARD, NFMA, true>(P.x, P.y, P.z, ca, cb, e.atol, e.eps, &gm);
            float ox, oy, oz;
            mantle_offset<NFMA>(gm, P.x, P.y, P.z, a.move_to_mantle != 0, ox, oy, oz);
            const int32_t id = (a.out_id || a.out_packed) ? a.ids[ci] : 0;
            e.win[row] = static_cast<int32_t>(ci);
            if (a.out_index && a.out_index != e.win) a.out_index[row] = static_cast<int32_t>(ci);
            if (a.out_id) a.out_id[row] = id;
            if (a.out_dist) a.out_dist[row] = gm.dist;
            if (a.out_offset) { a.out_offset[3 * row] = ox; a.out_offset[3 * row + 1] = oy; a.out_offset[3 * row + 2] = oz; }
            if (a.out_radius) a.out_radius[row] = cb.w;
            if (a.out_packed) a.out_packed[row] = make_float4(ox, oy, oz, __int_as_float(id));
            ++evals;
        }
        // ---- everything else: work lists
        const bool late = valid && !fast;
        if (__any_sync(0xffffffffu, late)) {
            const uint32_t sl = warp_append(late, &e.st->late_rows, lane, lt);
            if (late) a.late_rows[sl] = static_cast<uint32_t>(row);
            // outside the grid: straight to the tree search (non-finite coordinates: the exhaustive kernel); voxels without
            // any tile entry and no list to run: pending without an incumbent; the rest: exact kernel
            const uint32_t tile_total = inside ? a.tile_desc[mycode].z : 0u;
            const bool outside = late && !inside;
            const bool empty = late && inside && !lists && tile_total == 0u;
            const bool pend = outside || empty;
            const uint32_t sp = warp_append(pend, &e.st->pending, lane, lt);
            if (pend) {
                e.pend_idx[sp] = static_cast<int32_t>(row) | (outside ? OUTSIDE_BIT : 0);
                e.pend_keys[sp] = KEY_NONE;
                if (outside && !(fabsf(P.x) + fabsf(P.y) + fabsf(P.z) < 3.0e38f)) a.brute_slots[atomicAdd(&e.st->n_brute, 1u)] = sp;
            }
            const uint4 rec = make_uint4(static_cast<uint32_t>(row), mycode, __float_as_uint(up), 0u);
            const bool front = late && !pend && cert, back = late && !pend && !cert;
            const uint32_t sf = warp_append(front, &e.st->undecided_near, lane, lt);
            if (front) e.undecided[sf] = rec;
            const uint32_t sb = warp_append(back, &e.st->undecided_far, lane, lt);
            if (back) e.undecided[e.undecided_cap - 1u - sb] = rec;
        }
        __syncthreads();                 // the shared arrays are rewritten by the next block of rows
    }
    unsigned long long all_bounds = bounds, all_evals = evals;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        all_bounds += __shfl_xor_sync(0xffffffffu, all_bounds, o);
        all_evals += __shfl_xor_sync(0xffffffffu, all_evals, o);
    }
    if (lane == 0 && (all_bounds | all_evals)) {
        atomicAdd(&e.st->bound_tests, all_bounds);
        atomicAdd(&e.st->pairs_grid, all_evals);
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
    int32_t *pend_idx;             // DONE_BIT is set on the slots the ring search certifies (the tree search skips them)
    unsigned long long *pend_keys;
    int32_t *win;
    const uint32_t *tile_start, *tile_cnt;
    const float4 *tileAB;
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
                    const float4 ca = a.tileAB[2 * static_cast<size_t>(pos)], cb = a.tileAB[2 * static_cast<size_t>(pos) + 1];
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
            if (certified) {
                a.pend_idx[task] = ri | DONE_BIT;
                a.win[ri] = static_cast<int32_t>(key_index(key));
                atomicAdd(&a.st->ring_certified, 1u);
            }
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

// ------------------------------------------------------------------------------------------------
// host driver of the direct path
// ------------------------------------------------------------------------------------------------
// Small clouds take the direct path: the sorted path pays ~0.1 ms per call that does not depend on the number of points
// (scan over every voxel of the grid, a dozen launches), the direct path pays ~0.25 us per 1000 points for its uncoalesced
// tile reads.  Measured crossover on a randomly ordered cloud against 50k cylinders: ~500k points
// (profiles/r02_floor_direct_vs_sorted.json); TM_DIRECT=0/1 forces either.
static bool overlap_enabled() {
    static const bool on = [] { const char *e = getenv("TM_OVERLAP"); return !(e && e[0] == '0'); }();
    return on;
}

// TM_EXACT_LANES="far_warp,far_wide,near_wide" overrides the list lengths below which the exact kernel spends 32 / 8 / 8
// lanes on a walk (experiments)
static void exact_lane_thresholds(EvalArgs &ev) {
    static const std::array<uint32_t, 3> v = [] {
        std::array<uint32_t, 3> t{EXACT_FAR_WARP_BELOW, EXACT_FAR_WIDE_BELOW, EXACT_NEAR_WIDE_BELOW};
        if (const char *e = getenv("TM_EXACT_LANES")) {
            unsigned a = 0, b = 0, c = 0;
            if (sscanf(e, "%u,%u,%u", &a, &b, &c) == 3) t = {a, b, c};
        }
        return t;
    }();
    ev.far_warp_below = v[0]; ev.far_wide_below = v[1]; ev.near_wide_below = v[2];
}

static bool use_direct(const tm_handle *, const LabelArgs &a) {
    static const int forced = [] { const char *e = getenv("TM_DIRECT"); return e ? atoi(e) : -1; }();
    if (forced >= 0) return forced != 0;
    return a.n <= 500000;
}

static int label_direct(tm_handle *h, const LabelArgs &a, const GridDev &g, float slack) {
    cudaStream_t st = a.stream;
    const size_t n = static_cast<size_t>(a.n);
    TM_CUDA(h, h->pend_idx.ensure(sizeof(int32_t) * n));
    TM_CUDA(h, h->brute_slots.ensure(sizeof(uint32_t) * n));
    TM_CUDA(h, h->keys.ensure(sizeof(unsigned long long) * n));
    TM_CUDA(h, h->win.ensure(sizeof(int32_t) * n));
    TM_CUDA(h, h->undecided.ensure(sizeof(uint4) * n));
    TM_CUDA(h, h->late_rows.ensure(sizeof(uint32_t) * n));
    TM_CUDA(h, h->dstats.ensure(sizeof(DevStats) + 64));
    DevStats *dst = h->dstats.as<DevStats>();
    TM_CUDA(h, cudaMemsetAsync(h->dstats.p, 0, sizeof(DevStats) + 64, st));
    mark(h, 1, st); mark(h, 2, st); mark(h, 3, st);

    DirectArgs d;
    d.pts = a.pts; d.n = a.n; d.row_stride = a.row_stride;
    d.tile_desc = h->tile_desc.as<uint4>();
    EvalArgs &ev = d.ev;
    ev.items = nullptr; ev.warp_item = nullptr; ev.cursor = nullptr; ev.sorted = nullptr;
    ev.tileAB = h->tileAB.as<float4>(); ev.tileI = h->tileI.as<int32_t>();
    ev.recA = h->recA.as<float4>(); ev.recB = h->recB.as<float4>();
    ev.special = h->special.as<int32_t>(); ev.aligned = h->aligned.as<int32_t>(); ev.long_list = h->long_list.as<int32_t>();
    ev.n_special = h->n_special; ev.n_aligned = h->n_aligned; ev.n_long = h->n_long;
    ev.tileLB = h->tileLB.as<float>();
    ev.atol = a.prm.perp_atol; ev.eps = a.prm.norm_eps;
    ev.slack = slack; ev.near = h->near; ev.reach = h->reach;
    ev.amb = EST_ALLOWANCE * slack;
    if (const char *env = getenv("TM_AMB_FACTOR")) { const float v = static_cast<float>(atof(env)); if (v > 0.f && v <= 1.f) ev.amb = slack * v; }
    int32_t *win = a.out_index ? a.out_index : h->win.as<int32_t>();
    ev.win = win;
    ev.undecided = h->undecided.as<uint4>();
    ev.undecided_cap = static_cast<uint32_t>(n);
    ev.pend_idx = h->pend_idx.as<int32_t>();
    ev.pend_keys = h->keys.as<unsigned long long>();
    ev.st = dst;
    ev.pts = a.pts; ev.row_stride = a.row_stride; ev.tile_desc = h->tile_desc.as<uint4>();
    exact_lane_thresholds(ev);
    d.recAB = h->recAB.as<float4>();
    d.ids = h->ids.as<int32_t>();
    d.move_to_mantle = a.prm.move_to_mantle;
    d.out_index = a.out_index; d.out_id = a.out_id; d.out_dist = a.out_dist; d.out_offset = a.out_offset;
    d.out_radius = a.out_radius; d.out_packed = a.out_packed;
    d.late_rows = h->late_rows.as<uint32_t>();
    d.brute_slots = h->brute_slots.as<uint32_t>();
    const bool guard = a.prm.norm_eps > 0.f, nfma = a.prm.norm_fma != 0;
    const bool wide = a.prm.perp_atol > 2.f * ev.amb;
    if (guard) ev.n_aligned = 0;
    const int blocks = static_cast<int>(std::min<size_t>((n + DIRECT_THREADS - 1) / DIRECT_THREADS, static_cast<size_t>(h->sm_count) * 48));
#define TM_DIRECT_CASE(G, F)                                                                                       \
    do {                                                                                                           \
        if (wide) direct_kernel<G, F, true><<<blocks, DIRECT_THREADS, 0, st>>>(d, g);                              \
        else direct_kernel<G, F, false><<<blocks, DIRECT_THREADS, 0, st>>>(d, g);                                  \
    } while (0)
    if (guard) { if (nfma) TM_DIRECT_CASE(true, true); else TM_DIRECT_CASE(true, false); }
    else       { if (nfma) TM_DIRECT_CASE(false, true); else TM_DIRECT_CASE(false, false); }
#undef TM_DIRECT_CASE
    TM_KCHECK(h, st, "direct_kernel");
    h->stats.launches += 1;
    const int ev_blocks = h->sm_count * 4;
    if (guard) { if (nfma) exact_kernel<true, true, true><<<ev_blocks, EV_WARPS * 32, 0, st>>>(ev); else exact_kernel<true, false, true><<<ev_blocks, EV_WARPS * 32, 0, st>>>(ev); }
    else       { if (nfma) exact_kernel<false, true, true><<<ev_blocks, EV_WARPS * 32, 0, st>>>(ev); else exact_kernel<false, false, true><<<ev_blocks, EV_WARPS * 32, 0, st>>>(ev); }
    TM_KCHECK(h, st, "exact_kernel");
    h->stats.launches += 1;

    mark(h, 4, st);
    RingArgs rg;
    rg.pts = a.pts; rg.row_stride = a.row_stride;
    rg.pend_idx = h->pend_idx.as<int32_t>();
    rg.pend_keys = h->keys.as<unsigned long long>();
    rg.win = win;
    rg.tile_start = h->cyl_cell_start.as<uint32_t>(); rg.tile_cnt = h->cyl_cell_cnt.as<uint32_t>();
    rg.tileAB = h->tileAB.as<float4>(); rg.tileI = h->tileI.as<int32_t>();
    rg.atol = a.prm.perp_atol; rg.eps = a.prm.norm_eps;
    rg.st = dst;
    // one CTA per pending point; far fewer points than rows ever reach it, so small calls get a small grid
    const int rg_blocks = static_cast<int>(std::min<size_t>(static_cast<size_t>(h->sm_count) * (2048 / (RING_WARPS * 32)), std::max<size_t>(h->sm_count, n / 256)));
    if (guard) { if (nfma) ring_kernel<true, true><<<rg_blocks, RING_WARPS * 32, 0, st>>>(rg, g); else ring_kernel<true, false><<<rg_blocks, RING_WARPS * 32, 0, st>>>(rg, g); }
    else       { if (nfma) ring_kernel<false, true><<<rg_blocks, RING_WARPS * 32, 0, st>>>(rg, g); else ring_kernel<false, false><<<rg_blocks, RING_WARPS * 32, 0, st>>>(rg, g); }
    TM_KCHECK(h, st, "ring_kernel");
    h->stats.launches += 1;
    int rc = search_bvh(h, a, dst, win);
    if (rc != TM_OK) return rc;
    mark(h, 5, st);
    rc = finish_pending(h, a, dst, win, h->maxabs);
    if (rc != TM_OK) return rc;
    // outputs of the rows the direct kernel left open
    mark(h, 7, st);
    return finalize_rows(h, a, win, h->late_rows.as<uint32_t>(), &dst->late_rows);
}

int label_grid(tm_handle *h, const LabelArgs &a) {
    if (a.n == 0) return TM_OK;
    if (a.n >= 0x40000000LL) return fail(h, TM_ERR_INVALID, "tm_label_points: more than 2^30-1 points per call%s%s");
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
    if (use_direct(h, a)) return label_direct(h, a, g, slack);

    // scratch
    const size_t max_occ = std::min<size_t>(n, ncodes);
    const size_t max_wslots = (n / PTS_PER_LANE + max_occ) / 32 + 2;
    // dense clouds (hundreds of points per voxel) queue their atomics on the voxels' counters: split each counter into
    // 2 / 4 / 8 sub-cells.  The density is judged by the voxels that have a tile at all (known per table).
    int nsub = 1;
    {
        const double density = static_cast<double>(n) / std::max<uint32_t>(1u, h->voxels_with_tiles);
        if (density > 640.0) nsub = 8; else if (density > 320.0) nsub = 4; else if (density > 160.0) nsub = 2;
        if (const char *env = getenv("TM_SUBCELLS")) { const int v = atoi(env); if (v == 1 || v == 2 || v == 4 || v == 8) nsub = v; }
    }
    // small clouds do not queue on the voxel counters: dense 8-byte cells (4x less to clear and to scan); large ones pad
    // every cell to its own 32-byte sector
    const int pad = n <= 3000000 ? 1 : CELL_PAD;
    TM_CUDA(h, h->cells.ensure(sizeof(uint2) * pad * nsub * (static_cast<size_t>(ncodes) + 1)));
    TM_CUDA(h, h->sorted_pts.ensure(sizeof(float4) * n));
    TM_CUDA(h, h->items.ensure(sizeof(uint4) * 2 * (max_occ + 1)));
    TM_CUDA(h, h->warp_item.ensure(sizeof(uint32_t) * max_wslots));
    TM_CUDA(h, h->undecided.ensure(sizeof(uint4) * n));
    TM_CUDA(h, h->pend_idx.ensure(sizeof(int32_t) * n));
    TM_CUDA(h, h->brute_slots.ensure(sizeof(uint32_t) * n));
    TM_CUDA(h, h->keys.ensure(sizeof(unsigned long long) * n));
    if (!a.out_index) TM_CUDA(h, h->win.ensure(sizeof(int32_t) * n));
    TM_CUDA(h, h->dstats.ensure(sizeof(DevStats) + 64));
    DevStats *dst = h->dstats.as<DevStats>();
    unsigned int *cursor = reinterpret_cast<unsigned int *>(h->dstats.as<unsigned char>() + sizeof(DevStats) + 16);
    TM_CUDA(h, cudaMemsetAsync(h->dstats.p, 0, sizeof(DevStats) + 64, st));
    TM_CUDA(h, cudaMemsetAsync(h->cells.p, 0, sizeof(uint2) * pad * nsub * ncodes, st));

    const int pt_blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(h->sm_count) * 32));
    bin_count_kernel<<<pt_blocks, 256, 0, st>>>(a.pts, a.n, a.row_stride, g, nsub, pad, h->cells.as<uint2>(),
                                                h->pend_idx.as<int32_t>(), h->keys.as<unsigned long long>(),
                                                h->brute_slots.as<uint32_t>(), a.out_index ? a.out_index : h->win.as<int32_t>(), dst);
    TM_KCHECK(h, st, "bin_count_kernel");
    h->stats.launches += 1;
    mark(h, 1, st);
    int rc = run_scan(h, h->cells.as<uint32_t>(), ncodes, 1, h->cells.as<uint32_t>(), h->cyl_cell_start.as<uint32_t>(),
                      h->cyl_cell_cnt.as<uint32_t>(), h->cyl_cell_near.as<uint32_t>(), h->items.as<uint4>(),
                      h->warp_item.as<uint32_t>(), dst, st, nsub, pad);
    if (rc != TM_OK) return rc;
    TM_KCHECK(h, st, "scan kernels");
    h->stats.launches += 3;
    mark(h, 2, st);
    bin_scatter_kernel<<<pt_blocks, 256, 0, st>>>(a.pts, a.n, a.row_stride, g, nsub, pad, h->cells.as<uint2>(), h->sorted_pts.as<float4>());
    TM_KCHECK(h, st, "bin_scatter_kernel");
    h->stats.launches += 1;

    mark(h, 3, st);
    EvalArgs ev;
    ev.items = h->items.as<uint4>();
    ev.warp_item = h->warp_item.as<uint32_t>();
    ev.cursor = cursor;
    ev.sorted = h->sorted_pts.as<float4>();
    ev.tileAB = h->tileAB.as<float4>(); ev.tileI = h->tileI.as<int32_t>();
    ev.recA = h->recA.as<float4>(); ev.recB = h->recB.as<float4>();
    ev.special = h->special.as<int32_t>();
    ev.aligned = h->aligned.as<int32_t>();
    ev.long_list = h->long_list.as<int32_t>();
    ev.n_special = h->n_special; ev.n_aligned = h->n_aligned; ev.n_long = h->n_long;
    ev.atol = a.prm.perp_atol; ev.eps = a.prm.norm_eps;
    ev.tileLB = h->tileLB.as<float>();
    ev.slack = slack; ev.near = h->near; ev.reach = h->reach;
    // allowance of the estimate-vs-reference comparison: a quarter of the rounding slack (the reference's fp32 distance is
    // within 3 % of the slack of the closed form, tests/test_estimate_bound.py; TM_AMB_FACTOR overrides, never above the slack)
    ev.amb = EST_ALLOWANCE * slack;
    if (const char *env = getenv("TM_AMB_FACTOR")) { const float v = static_cast<float>(atof(env)); if (v > 0.f && v <= 1.f) ev.amb = slack * v; }
    int32_t *win = a.out_index ? a.out_index : h->win.as<int32_t>();       // the caller's index array doubles as the scatter target
    ev.win = win;
    ev.undecided = h->undecided.as<uint4>();
    ev.undecided_cap = static_cast<uint32_t>(n);
    ev.pend_idx = h->pend_idx.as<int32_t>();
    ev.pend_keys = h->keys.as<unsigned long long>();
    ev.st = dst;
    ev.pts = a.pts; ev.row_stride = a.row_stride; ev.tile_desc = h->tile_desc.as<uint4>();
    exact_lane_thresholds(ev);
    const bool guard = a.prm.norm_eps > 0.f, nfma = a.prm.norm_fma != 0;
    const bool wide = a.prm.perp_atol > 2.f * ev.amb;
    if (guard) ev.n_aligned = 0;              // variant B never yields NaN on an axis line
    const int ev_blocks = h->sm_count * 4;
    const size_t stage_bytes = EV_STAGE_BYTES * EV_WARPS;
    if (wide) {
        TM_CUDA(h, cudaFuncSetAttribute(evaluate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(stage_bytes)));
        evaluate_kernel<true><<<h->sm_count * EV_BLOCKS_PER_SM, EV_WARPS * 32, stage_bytes, st>>>(ev);
    } else {
        TM_CUDA(h, cudaFuncSetAttribute(evaluate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(stage_bytes)));
        evaluate_kernel<false><<<h->sm_count * EV_BLOCKS_PER_SM, EV_WARPS * 32, stage_bytes, st>>>(ev);
    }
    TM_KCHECK(h, st, "evaluate_kernel");
    h->stats.launches += 1;
    if (guard) { if (nfma) exact_kernel<true, true, false><<<ev_blocks, EV_WARPS * 32, 0, st>>>(ev); else exact_kernel<true, false, false><<<ev_blocks, EV_WARPS * 32, 0, st>>>(ev); }
    else       { if (nfma) exact_kernel<false, true, false><<<ev_blocks, EV_WARPS * 32, 0, st>>>(ev); else exact_kernel<false, false, false><<<ev_blocks, EV_WARPS * 32, 0, st>>>(ev); }
    TM_KCHECK(h, st, "exact_kernel");
    h->stats.launches += 1;

    // still uncertified at D_max (beyond the far part of their own tile): a handful of points -> ring search, one CTA per
    // point; many (clutter), or outside the grid -> per-point descent of the bounding-volume hierarchy.
    // These kernels are latency chains over a handful of points while the epilogue streams over every row, so the two run
    // side by side: the stragglers are searched on the handle's high-priority side stream, the epilogue of ALL rows runs on
    // the caller's stream as soon as the exact kernel is done (every pending row already holds a valid provisional winner),
    // and once both are through a short list epilogue rewrites the outputs of exactly the pending rows.  (Serial when the phases are
    // being timed, so that the per-phase numbers keep their meaning.)
    const bool overlap = !h->profiling && overlap_enabled();
    cudaStream_t ss = st;                     // stream of the straggler kernels
    LabelArgs as = a;
    if (overlap) {
        if (!h->side_stream) {
            int lo_prio = 0, hi_prio = 0;
            TM_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
            TM_CUDA(h, cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, hi_prio));     // few CTAs, long chains: first in line
        }
        if (!h->fork_ev) TM_CUDA(h, cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
        if (!h->join_ev) TM_CUDA(h, cudaEventCreateWithFlags(&h->join_ev, cudaEventDisableTiming));
        TM_CUDA(h, cudaEventRecord(h->fork_ev, st));
        TM_CUDA(h, cudaStreamWaitEvent(h->side_stream, h->fork_ev, 0));
        ss = h->side_stream;
        as.stream = ss;
    }
    mark(h, 4, st);
    RingArgs rg;
    rg.pts = a.pts; rg.row_stride = a.row_stride;
    rg.pend_idx = h->pend_idx.as<int32_t>();
    rg.pend_keys = h->keys.as<unsigned long long>();
    rg.win = win;
    rg.tile_start = h->cyl_cell_start.as<uint32_t>(); rg.tile_cnt = h->cyl_cell_cnt.as<uint32_t>();
    rg.tileAB = h->tileAB.as<float4>(); rg.tileI = h->tileI.as<int32_t>();
    rg.atol = a.prm.perp_atol; rg.eps = a.prm.norm_eps;
    rg.st = dst;
    // one CTA per pending point; far fewer points than rows ever reach it, so small calls get a small grid
    const int rg_blocks = static_cast<int>(std::min<size_t>(static_cast<size_t>(h->sm_count) * (2048 / (RING_WARPS * 32)), std::max<size_t>(h->sm_count, n / 256)));
    if (guard) { if (nfma) ring_kernel<true, true><<<rg_blocks, RING_WARPS * 32, 0, ss>>>(rg, g); else ring_kernel<true, false><<<rg_blocks, RING_WARPS * 32, 0, ss>>>(rg, g); }
    else       { if (nfma) ring_kernel<false, true><<<rg_blocks, RING_WARPS * 32, 0, ss>>>(rg, g); else ring_kernel<false, false><<<rg_blocks, RING_WARPS * 32, 0, ss>>>(rg, g); }
    TM_KCHECK(h, ss, "ring_kernel");
    h->stats.launches += 1;
    rc = search_bvh(h, as, dst, win);
    if (rc != TM_OK) return rc;

    // exhaustive search for non-finite points, then the winning rows of every pending point
    mark(h, 5, st);
    rc = finish_pending(h, as, dst, win, h->maxabs);
    if (rc != TM_OK) return rc;

    // winner-only epilogue: label + offset, streaming
    mark(h, 7, st);
    if (overlap) {
        TM_CUDA(h, cudaEventRecord(h->join_ev, ss));
        rc = finalize_rows(h, a, win, nullptr, nullptr);                       // every row, on the caller's stream
        if (rc != TM_OK) return rc;
        TM_CUDA(h, cudaStreamWaitEvent(st, h->join_ev, 0));
        return finalize_rows(h, a, win, reinterpret_cast<const uint32_t *>(h->pend_idx.as<int32_t>()), &dst->pending);
    }
    return finalize_rows(h, a, win, nullptr, nullptr);
}

}  // namespace tmn

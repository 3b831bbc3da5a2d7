// Voxel-grid path: exact nearest-cylinder search with spatial pruning.
//
//   cylinders  --(once per table)-->  solid AABBs binned into a uniform voxel grid (CSR home lists)
//   points     --bin-->  counting sort by voxel id (brick-Morton order)  --> contiguous per-voxel runs
//   per occupied voxel V:  R(V) = min_c hi(V,c)   an upper bound on the best distance of ANY point in V
//                          tile(V) = { c : lb(c,V) <= R(V) }   packed contiguously (float4 A | float4 B | idx)
//   evaluate:  one warp per (voxel, <=64 points) item; the tile is staged into shared memory with
//              bulk async copies (double buffered, next item's tile prefetched), each lane scans it for
//              its two points with the reference arithmetic, 64-bit (distance,index) keys reproduce
//              torch.argmin, and the winner-only epilogue writes label + offset at the original row.
//
// Exactness: for every point p and cylinder c the reference distance satisfies
//   dist_ref(p,c) >= dist(p, Solid(c)) >= lb(c,V)     and     min_c dist_ref(p,c) <= hi(V,c') for all c',
// so the true argmin of every point of V is inside tile(V); `slack` absorbs fp32 rounding of the
// reference pipeline.  Points the grid cannot serve (outside it, non-finite, voxels whose tile would
// be too large) go to the exhaustive kernel, so results never depend on the pruning.
#include <algorithm>
#include <cmath>

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
    float delta;           // half diagonal of a voxel + slack
};

__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {      // 3 bits -> every third bit
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4);
}
__host__ __device__ __forceinline__ uint32_t compact3(uint32_t m) {
    return (m & 1u) | ((m >> 2) & 2u) | ((m >> 4) & 4u);
}
__host__ __device__ __forceinline__ uint32_t voxel_code(const GridDev &g, int x, int y, int z) {
    const uint32_t brick = (static_cast<uint32_t>(z >> 3) * g.bny + static_cast<uint32_t>(y >> 3)) * g.bnx +
                           static_cast<uint32_t>(x >> 3);
    return (brick << 9) | spread3(x & 7) | (spread3(y & 7) << 1) | (spread3(z & 7) << 2);
}
__host__ __device__ __forceinline__ void voxel_decode(const GridDev &g, uint32_t code, int &x, int &y, int &z) {
    const uint32_t brick = code >> 9, m = code & 511u;
    const uint32_t bx = brick % g.bnx, by = (brick / g.bnx) % g.bny, bz = brick / (g.bnx * g.bny);
    x = static_cast<int>(bx * 8 + compact3(m));
    y = static_cast<int>(by * 8 + compact3(m >> 1));
    z = static_cast<int>(bz * 8 + compact3(m >> 2));
}

static GridDev to_dev(const GridDesc &d, float slack) {
    GridDev g;
    g.ox = d.ox; g.oy = d.oy; g.oz = d.oz; g.h = d.h; g.inv_h = d.inv_h;
    g.nx = d.nx; g.ny = d.ny; g.nz = d.nz;
    g.bnx = (d.nx + 7) / 8; g.bny = (d.ny + 7) / 8; g.bnz = (d.nz + 7) / 8;
    g.slack = slack;
    g.delta = 0.8660254f * d.h * 1.0001f + slack;
    return g;
}

__device__ __forceinline__ int cell_coord(float p, float o, float inv_h, int n) {
    const float f = (p - o) * inv_h;
    int c = static_cast<int>(floorf(f));
    return min(max(c, 0), n - 1);
}

// ------------------------------------------------------------------------------------------------
// cylinder side: home lists
// ------------------------------------------------------------------------------------------------
constexpr int LONG_CELLS = 512;      // AABBs spanning more voxels than this go to the "long" list

// count (pass 0) or fill (pass 1) the home lists: voxel -> cylinders whose solid AABB overlaps it
__global__ void cyl_register_kernel(const float4 *__restrict__ boxlo, const float4 *__restrict__ boxhi, int m, GridDev g,
                                    int pass, uint32_t *__restrict__ cell_counter, const uint32_t *__restrict__ cell_start,
                                    int32_t *__restrict__ cell_list, int32_t *__restrict__ long_list,
                                    unsigned int *__restrict__ n_long) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const float4 lo = boxlo[c], hi = boxhi[c];
    if (!(lo.w == 0.f)) return;                       // w != 0 marks special (non-finite / non-unit) cylinders
    const int x0 = cell_coord(lo.x, g.ox, g.inv_h, g.nx), x1 = cell_coord(hi.x, g.ox, g.inv_h, g.nx);
    const int y0 = cell_coord(lo.y, g.oy, g.inv_h, g.ny), y1 = cell_coord(hi.y, g.oy, g.inv_h, g.ny);
    const int z0 = cell_coord(lo.z, g.oz, g.inv_h, g.nz), z1 = cell_coord(hi.z, g.oz, g.inv_h, g.nz);
    const long long cells = static_cast<long long>(x1 - x0 + 1) * (y1 - y0 + 1) * (z1 - z0 + 1);
    if (cells > LONG_CELLS) {
        if (pass == 0) { const unsigned int s = atomicAdd(n_long, 1u); long_list[s] = c; }
        return;
    }
    for (int z = z0; z <= z1; ++z)
        for (int y = y0; y <= y1; ++y)
            for (int x = x0; x <= x1; ++x) {
                const uint32_t code = voxel_code(g, x, y, z);
                if (pass == 0) atomicAdd(&cell_counter[code], 1u);
                else { const uint32_t s = atomicAdd(&cell_counter[code], 1u); cell_list[cell_start[code] + s] = c; }
            }
}

// ------------------------------------------------------------------------------------------------
// generic 3-channel exclusive scan over the voxel arrays (points, occupied flags, work items)
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;     // 4096 voxels per block
constexpr int PTS_PER_ITEM = 64;

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
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t *__restrict__ count, uint32_t ncodes,
                                                                   Tri *__restrict__ block_sums) {
    __shared__ Tri total;
    const uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    Tri v{0, 0, 0};
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < ncodes) v = tri_add(v, tri_of_count(count[base + i]));
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
        st->points_grid = carry.a;
        st->voxels_occupied = carry.b;
        st->work_items = carry.c;
    }
}

// phase C: final offsets.  mode 0 (cylinder lists): start[code] only.  mode 1 (points): also the
// compact list of occupied voxels (in id order) and the first work item of each.
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t *__restrict__ count, uint32_t ncodes,
                                                                  const Tri *__restrict__ block_sums, int mode,
                                                                  uint32_t *__restrict__ start, uint32_t *__restrict__ occ_cells,
                                                                  uint32_t *__restrict__ occ_item_start) {
    const uint32_t base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
    uint32_t cnt[SCAN_ITEMS];
    Tri v{0, 0, 0};
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        cnt[i] = (base + i < ncodes) ? count[base + i] : 0u;
        v = tri_add(v, tri_of_count(cnt[i]));
    }
    Tri run = tri_add(block_exclusive_scan(v, nullptr), block_sums[blockIdx.x]);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < ncodes) {
            start[base + i] = run.a;
            if (mode == 1 && cnt[i]) { occ_cells[run.b] = base + i; occ_item_start[run.b] = run.c; }
        }
        run = tri_add(run, tri_of_count(cnt[i]));
    }
    if (mode == 0 && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) start[ncodes] = run.a;
}

static int run_scan(tm_handle *h, const uint32_t *count, uint32_t ncodes, int mode, uint32_t *start, uint32_t *occ_cells,
                    uint32_t *occ_item_start, DevStats *st, cudaStream_t stream) {
    const uint32_t nblocks = (ncodes + SCAN_BLOCK - 1) / SCAN_BLOCK;
    TM_CUDA(h, h->block_sums.ensure(sizeof(Tri) * nblocks));
    Tri *bs = h->block_sums.as<Tri>();
    scan_reduce_kernel<<<nblocks, SCAN_THREADS, 0, stream>>>(count, ncodes, bs);
    scan_blocks_kernel<<<1, SCAN_THREADS, 0, stream>>>(bs, nblocks, st);
    scan_apply_kernel<<<nblocks, SCAN_THREADS, 0, stream>>>(count, ncodes, bs, mode, start, occ_cells, occ_item_start);
    TM_CUDA(h, cudaGetLastError());
    return TM_OK;
}

// ------------------------------------------------------------------------------------------------
// host: size the grid from the cylinders' bounding box and build the home lists
// ------------------------------------------------------------------------------------------------
static inline float ordered_to_float(int k) {
    int i = k >= 0 ? k : k ^ 0x7fffffff;
    float f;
    memcpy(&f, &i, 4);
    return f;
}

constexpr float GRID_MARGIN_M = 1.5f;               // points farther than this outside the QSM box are outliers
constexpr uint32_t MAX_CODES = 1u << 26;

float auto_cell_size(const tm_handle *, int64_t) { return 0.25f; }

int build_cylinder_index(tm_handle *h, float cell_size, cudaStream_t stream) {
    // global bounding box of the regular cylinders' AABBs (written by pack_kernel as ordered ints)
    int host_box[8];
    TM_CUDA(h, cudaMemcpyAsync(host_box, h->bbox.p, sizeof(host_box), cudaMemcpyDeviceToHost, stream));
    TM_CUDA(h, cudaStreamSynchronize(stream));
    h->n_special = static_cast<uint32_t>(host_box[6]);
    const uint32_t n_regular = static_cast<uint32_t>(host_box[7]);
    h->have_grid = false;
    if (n_regular == 0) {               // nothing to index: every point goes to the exhaustive kernel
        h->grid = GridDesc{};
        h->grid_cell = cell_size;
        h->have_grid = true;
        h->n_listed = 0;
        return TM_OK;
    }
    float lo[3], hi[3];
    for (int k = 0; k < 3; ++k) { lo[k] = ordered_to_float(host_box[k]); hi[k] = ordered_to_float(host_box[3 + k]); }
    float hcell = cell_size;
    GridDesc d{};
    for (;;) {
        double codes = 1;
        int n[3];
        for (int k = 0; k < 3; ++k) {
            n[k] = static_cast<int>(std::ceil((static_cast<double>(hi[k]) - lo[k] + 2.0 * GRID_MARGIN_M) / hcell));
            n[k] = std::max(n[k], 1);
            codes *= ((n[k] + 7) / 8) * 8.0;
        }
        if (codes <= MAX_CODES) {
            d.ox = lo[0] - GRID_MARGIN_M; d.oy = lo[1] - GRID_MARGIN_M; d.oz = lo[2] - GRID_MARGIN_M;
            d.h = hcell; d.inv_h = 1.0f / hcell;
            d.nx = n[0]; d.ny = n[1]; d.nz = n[2];
            d.ncell_codes = static_cast<uint32_t>(codes);
            break;
        }
        hcell *= 1.5f;                  // coarsen until the voxel arrays fit
    }
    d.bx = d.by = d.bz = 0;
    h->grid = d;
    h->grid_cell = cell_size;
    float maxabs = 0.f;
    for (int k = 0; k < 3; ++k) maxabs = std::max(maxabs, std::max(std::fabs(lo[k]), std::fabs(hi[k])) + GRID_MARGIN_M);
    const float slack = 32.f * 5.96e-8f * maxabs + 2e-6f;
    const GridDev g = to_dev(d, slack);

    const int m = static_cast<int>(h->m);
    const uint32_t ncodes = d.ncell_codes;
    TM_CUDA(h, h->cell_count.ensure(sizeof(uint32_t) * (static_cast<size_t>(ncodes) + 1)));
    TM_CUDA(h, h->cyl_cell_start.ensure(sizeof(uint32_t) * (static_cast<size_t>(ncodes) + 1)));
    TM_CUDA(h, h->long_list.ensure(sizeof(int32_t) * static_cast<size_t>(m)));
    TM_CUDA(h, h->dstats.ensure(sizeof(DevStats) + 64));
    uint32_t *counter = h->cell_count.as<uint32_t>();
    unsigned int *d_nlong = reinterpret_cast<unsigned int *>(h->dstats.as<unsigned char>() + sizeof(DevStats));
    TM_CUDA(h, cudaMemsetAsync(counter, 0, sizeof(uint32_t) * ncodes, stream));
    TM_CUDA(h, cudaMemsetAsync(d_nlong, 0, sizeof(unsigned int), stream));
    const int blocks = (m + 127) / 128;
    cyl_register_kernel<<<blocks, 128, 0, stream>>>(h->boxlo.as<float4>(), h->boxhi.as<float4>(), m, g, 0, counter, nullptr,
                                                    nullptr, h->long_list.as<int32_t>(), d_nlong);
    TM_CUDA(h, cudaGetLastError());
    int rc = run_scan(h, counter, ncodes, 0, h->cyl_cell_start.as<uint32_t>(), nullptr, nullptr, nullptr, stream);
    if (rc != TM_OK) return rc;
    uint32_t total = 0, nlong = 0;
    TM_CUDA(h, cudaMemcpyAsync(&total, h->cyl_cell_start.as<uint32_t>() + ncodes, 4, cudaMemcpyDeviceToHost, stream));
    TM_CUDA(h, cudaMemcpyAsync(&nlong, d_nlong, 4, cudaMemcpyDeviceToHost, stream));
    TM_CUDA(h, cudaStreamSynchronize(stream));
    h->n_long = nlong;
    h->cyl_list_len = total;
    h->n_listed = n_regular - nlong;
    TM_CUDA(h, h->cyl_cell_list.ensure(sizeof(int32_t) * (static_cast<size_t>(total) + 1)));
    TM_CUDA(h, cudaMemsetAsync(counter, 0, sizeof(uint32_t) * ncodes, stream));
    cyl_register_kernel<<<blocks, 128, 0, stream>>>(h->boxlo.as<float4>(), h->boxhi.as<float4>(), m, g, 1, counter,
                                                    h->cyl_cell_start.as<uint32_t>(), h->cyl_cell_list.as<int32_t>(),
                                                    h->long_list.as<int32_t>(), d_nlong);
    TM_CUDA(h, cudaGetLastError());
    h->have_grid = true;
    return TM_OK;
}

// ------------------------------------------------------------------------------------------------
// point side: counting sort by voxel id
// ------------------------------------------------------------------------------------------------
constexpr uint32_t NO_CELL = 0xFFFFFFFFu;

__global__ void __launch_bounds__(256) bin_count_kernel(const float *__restrict__ pts, int64_t n, int64_t row_stride, GridDev g,
                                                        uint32_t *__restrict__ cell_count, uint32_t *__restrict__ pt_cell,
                                                        uint32_t *__restrict__ pt_rank, int32_t *__restrict__ outlier_idx,
                                                        unsigned long long *__restrict__ outlier_keys,
                                                        DevStats *__restrict__ st) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float *p = pts + i * row_stride;
        const float fx = (p[0] - g.ox) * g.inv_h, fy = (p[1] - g.oy) * g.inv_h, fz = (p[2] - g.oz) * g.inv_h;
        // NaN / Inf fail these comparisons and become outliers
        const bool inside = fx >= 0.f && fx < static_cast<float>(g.nx) && fy >= 0.f && fy < static_cast<float>(g.ny) &&
                            fz >= 0.f && fz < static_cast<float>(g.nz);
        if (inside) {
            const uint32_t code = voxel_code(g, static_cast<int>(fx), static_cast<int>(fy), static_cast<int>(fz));
            pt_cell[i] = code;
            pt_rank[i] = atomicAdd(&cell_count[code], 1u);
        } else {
            pt_cell[i] = NO_CELL;
            const unsigned int s = atomicAdd(&st->outliers, 1u);
            outlier_idx[s] = static_cast<int32_t>(i);
            outlier_keys[s] = KEY_NONE;
        }
    }
}

__global__ void __launch_bounds__(256) bin_scatter_kernel(const float *__restrict__ pts, int64_t n, int64_t row_stride,
                                                          const uint32_t *__restrict__ pt_cell,
                                                          const uint32_t *__restrict__ pt_rank,
                                                          const uint32_t *__restrict__ cell_start,
                                                          float4 *__restrict__ sorted) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint32_t code = pt_cell[i];
        if (code == NO_CELL) continue;
        const float *p = pts + i * row_stride;
        const uint32_t pos = cell_start[code] + pt_rank[i];
        sorted[pos] = make_float4(p[0], p[1], p[2], __int_as_float(static_cast<int>(i)));
    }
}

// ------------------------------------------------------------------------------------------------
// per-voxel bounds
// ------------------------------------------------------------------------------------------------
// hi(V,c): upper bound on dist_ref(p,c) for every p within `delta` of the voxel centre.
//   slab:  dist = sqrt((rho - r)^2 + d^2);  cap: sqrt(d^2 + max(rho - r, 0)^2)   
//   d and rho are 1-Lipschitz in p  =>  dist <= sqrt((|d_c| + delta)^2 + (|rho_c - r| + delta)^2)
__device__ __forceinline__ float bound_hi(const float4 A, const float4 B, float cx, float cy, float cz, float delta) {
    const float vx = cx - A.x, vy = cy - A.y, vz = cz - A.z;
    const float t = vx * B.x + vy * B.y + vz * B.z;
    const float tc = fminf(fmaxf(t, 0.f), A.w);
    const float wx = vx - tc * B.x, wy = vy - tc * B.y, wz = vz - tc * B.z;       // centre - clamped foot point
    const float d = wx * B.x + wy * B.y + wz * B.z;                              // axial overshoot
    const float rx = wx - d * B.x, ry = wy - d * B.y, rz = wz - d * B.z;
    const float rho = sqrtf(rx * rx + ry * ry + rz * rz);
    const float a = fabsf(rho - B.w) + delta;
    const float e = fabsf(d) + delta;
    return sqrtf(a * a + e * e);
}

// lb(c,V): lower bound on dist(p, Solid(c)) for every p in the voxel box (Solid uses |r|).
__device__ __forceinline__ float bound_lo(const float4 A, const float4 B, const float4 blo, const float4 bhi, float vlx,
                                          float vly, float vlz, float vhx, float vhy, float vhz, float cx, float cy,
                                          float cz, float delta) {
    const float gx = fmaxf(fmaxf(blo.x - vhx, vlx - bhi.x), 0.f);
    const float gy = fmaxf(fmaxf(blo.y - vhy, vly - bhi.y), 0.f);
    const float gz = fmaxf(fmaxf(blo.z - vhz, vlz - bhi.z), 0.f);
    const float lb_box = sqrtf(gx * gx + gy * gy + gz * gz);
    const float vx = cx - A.x, vy = cy - A.y, vz = cz - A.z;
    const float t = fminf(fmaxf(vx * B.x + vy * B.y + vz * B.z, 0.f), A.w);
    const float wx = vx - t * B.x, wy = vy - t * B.y, wz = vz - t * B.z;
    const float lb_capsule = sqrtf(wx * wx + wy * wy + wz * wz) - fabsf(B.w) - delta;
    return fmaxf(lb_box, lb_capsule);
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------
// tile build: one warp per occupied voxel
// ------------------------------------------------------------------------------------------------
constexpr int TB_WARPS = 8;
constexpr int TILE_MAX = 1024;        // candidates per voxel; larger tiles send the voxel to the exhaustive kernel
constexpr int RING_MAX = 8;           // search radius (in voxels) for the upper bound
constexpr int GATHER_MAX = 10;        // gather radius (in voxels)

struct TileBuildArgs {
    const uint32_t *occ_cells;
    const uint32_t *occ_item_start;
    const uint32_t *pt_count;          // per voxel code
    const uint32_t *pt_start;          // per voxel code
    const uint32_t *cyl_start;         // CSR of the home lists
    const int32_t *cyl_list;
    const int32_t *long_list;
    uint32_t n_long;
    const float4 *recA, *recB, *boxlo, *boxhi;
    const float4 *sorted;
    float4 *tileA, *tileB;
    int32_t *tileI;
    uint32_t pool_entries;
    uint4 *items;
    int32_t *outlier_idx;
    unsigned long long *outlier_keys;
    DevStats *st;
};

__global__ void __launch_bounds__(TB_WARPS * 32) tile_build_kernel(TileBuildArgs a, GridDev g) {
    __shared__ int32_t s_list[TB_WARPS][TILE_MAX];
    __shared__ unsigned int s_count[TB_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_occ = a.st->voxels_occupied;
    for (uint32_t k = blockIdx.x * TB_WARPS + warp; k < n_occ; k += gridDim.x * TB_WARPS) {
        const uint32_t code = a.occ_cells[k];
        const uint32_t pcount = a.pt_count[code], pstart = a.pt_start[code], item0 = a.occ_item_start[k];
        int cx, cy, cz;
        voxel_decode(g, code, cx, cy, cz);
        const float vlx = g.ox + cx * g.h - g.slack, vly = g.oy + cy * g.h - g.slack, vlz = g.oz + cz * g.h - g.slack;
        const float vhx = g.ox + (cx + 1) * g.h + g.slack, vhy = g.oy + (cy + 1) * g.h + g.slack,
                    vhz = g.oz + (cz + 1) * g.h + g.slack;
        const float pcx = g.ox + (cx + 0.5f) * g.h, pcy = g.oy + (cy + 0.5f) * g.h, pcz = g.oz + (cz + 0.5f) * g.h;

        // ---- phase 1: R(V) = min over nearby cylinders of hi(V,c), expanding shells of voxels ----
        float R = __int_as_float(0x7f800000);
        for (uint32_t e = lane; e < a.n_long; e += 32) {
            const int c = a.long_list[e];
            R = fminf(R, bound_hi(a.recA[c], a.recB[c], pcx, pcy, pcz, g.delta));
        }
        R = warp_min(R);
        bool found = false;
        for (int ring = 0; ring <= RING_MAX; ++ring) {
            const int side = 2 * ring + 1, total = side * side * side;
            for (int idx = lane; idx < total; idx += 32) {
                const int dx = idx % side - ring, dy = (idx / side) % side - ring, dz = idx / (side * side) - ring;
                if (max(max(abs(dx), abs(dy)), abs(dz)) != ring) continue;
                const int x = cx + dx, y = cy + dy, z = cz + dz;
                if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz) continue;
                const uint32_t wc = voxel_code(g, x, y, z);
                const uint32_t e0 = a.cyl_start[wc], e1 = a.cyl_start[wc + 1];
                for (uint32_t e = e0; e < e1; ++e) {
                    const int c = a.cyl_list[e];
                    R = fminf(R, bound_hi(a.recA[c], a.recB[c], pcx, pcy, pcz, g.delta));
                }
            }
            R = warp_min(R);
            // cylinders not registered within `ring` shells keep their solid outside the cube of half
            // width (ring + 0.5) h around the centre, so their hi() cannot undercut that
            if (R <= (ring + 0.5f) * g.h - g.slack) { found = true; break; }
        }
        if (!found && R < __int_as_float(0x7f800000)) found = R <= (RING_MAX + 0.5f) * g.h;   // bound still valid, just loose
        R = R * 1.00001f + g.slack;

        // ---- phase 2: gather every cylinder whose solid may come within R of the voxel box ----
        if (lane == 0) s_count[warp] = 0;
        __syncwarp();
        bool brute = !found;
        if (!brute) {
            const int K = static_cast<int>(floorf(R * g.inv_h)) + 1;
            if (K > GATHER_MAX) brute = true;
            else {
                const int x0 = max(cx - K, 0), x1 = min(cx + K, g.nx - 1);
                const int y0 = max(cy - K, 0), y1 = min(cy + K, g.ny - 1);
                const int z0 = max(cz - K, 0), z1 = min(cz + K, g.nz - 1);
                const int sx = x1 - x0 + 1, sy = y1 - y0 + 1, sz = z1 - z0 + 1;
                const int total = sx * sy * sz;
                for (int idx = lane; idx < total; idx += 32) {
                    const int x = x0 + idx % sx, y = y0 + (idx / sx) % sy, z = z0 + idx / (sx * sy);
                    const uint32_t wc = voxel_code(g, x, y, z);
                    const uint32_t e0 = a.cyl_start[wc], e1 = a.cyl_start[wc + 1];
                    for (uint32_t e = e0; e < e1; ++e) {
                        const int c = a.cyl_list[e];
                        const float4 blo = a.boxlo[c], bhi = a.boxhi[c];
                        // a cylinder is registered in every voxel of its AABB range: take it only at the
                        // first voxel of that range inside the search window
                        const int fxc = max(cell_coord(blo.x, g.ox, g.inv_h, g.nx), x0);
                        const int fyc = max(cell_coord(blo.y, g.oy, g.inv_h, g.ny), y0);
                        const int fzc = max(cell_coord(blo.z, g.oz, g.inv_h, g.nz), z0);
                        if (fxc != x || fyc != y || fzc != z) continue;
                        const float lb = bound_lo(a.recA[c], a.recB[c], blo, bhi, vlx, vly, vlz, vhx, vhy, vhz, pcx, pcy, pcz,
                                                  g.delta);
                        if (lb <= R) {
                            const unsigned int s = atomicAdd(&s_count[warp], 1u);
                            if (s < TILE_MAX) s_list[warp][s] = c;
                        }
                    }
                }
                for (uint32_t e = lane; e < a.n_long; e += 32) {
                    const int c = a.long_list[e];
                    const float lb = bound_lo(a.recA[c], a.recB[c], a.boxlo[c], a.boxhi[c], vlx, vly, vlz, vhx, vhy, vhz, pcx,
                                              pcy, pcz, g.delta);
                    if (lb <= R) {
                        const unsigned int s = atomicAdd(&s_count[warp], 1u);
                        if (s < TILE_MAX) s_list[warp][s] = c;
                    }
                }
            }
        }
        __syncwarp();
        const uint32_t cnt = s_count[warp];
        if (cnt > TILE_MAX || cnt == 0) brute = true;

        // ---- phase 3: pack the tile and emit the work items ----
        uint32_t off = 0;
        if (!brute) {
            const uint32_t cnt4 = (cnt + 3u) & ~3u;
            if (lane == 0) off = atomicAdd(&a.st->tile_pool_used, cnt4);
            off = __shfl_sync(0xffffffffu, off, 0);
            if (off + cnt4 > a.pool_entries) brute = true;
        }
        const uint32_t n_items = (pcount + PTS_PER_ITEM - 1) / PTS_PER_ITEM;
        if (!brute) {
            for (uint32_t j = lane; j < cnt; j += 32) {
                const int c = s_list[warp][j];
                a.tileA[off + j] = a.recA[c];
                a.tileB[off + j] = a.recB[c];
                a.tileI[off + j] = c;
            }
            for (uint32_t t = lane; t < n_items; t += 32)
                a.items[item0 + t] = make_uint4(off, cnt, pstart + t * PTS_PER_ITEM, min(PTS_PER_ITEM, pcount - t * PTS_PER_ITEM));
            if (lane == 0) atomicAdd(&a.st->tile_entries, static_cast<unsigned long long>(cnt));
        } else {
            for (uint32_t t = lane; t < n_items; t += 32) a.items[item0 + t] = make_uint4(0, 0, 0, 0);
            unsigned int base = 0;
            if (lane == 0) { base = atomicAdd(&a.st->outliers, pcount); atomicAdd(&a.st->voxels_brute, 1u); }
            base = __shfl_sync(0xffffffffu, base, 0);
            for (uint32_t i = lane; i < pcount; i += 32) {
                a.outlier_idx[base + i] = __float_as_int(a.sorted[pstart + i].w);
                a.outlier_keys[base + i] = KEY_NONE;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// evaluate: persistent warps over work items, TMA-staged candidate tiles, fused label + offset write
// ------------------------------------------------------------------------------------------------
constexpr int EV_WARPS = 8;
constexpr int EV_CHUNK = 128;         // candidates per stage: 128 * (16 + 16 + 4) B = 4.5 KB

struct __align__(128) WarpStage {
    float4 A[2][EV_CHUNK];
    float4 B[2][EV_CHUNK];
    int32_t I[2][EV_CHUNK];
};

struct EvalArgs {
    const uint4 *items;
    unsigned int *cursor;
    const float4 *sorted;
    const float4 *tileA, *tileB;
    const int32_t *tileI;
    const float4 *recA, *recB;
    const int32_t *ids;
    const int32_t *special;
    uint32_t n_special;
    float atol, eps;
    int move_to_mantle;
    int32_t *out_index, *out_id;
    float *out_dist, *out_offset, *out_radius;
    DevStats *st;
};

template <bool GUARD, bool NFMA>
__global__ void __launch_bounds__(EV_WARPS * 32) evaluate_kernel(EvalArgs a) {
    extern __shared__ __align__(128) unsigned char ev_smem[];
    WarpStage *stages = reinterpret_cast<WarpStage *>(ev_smem);
    uint64_t *bars = reinterpret_cast<uint64_t *>(ev_smem + sizeof(WarpStage) * EV_WARPS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
    unsigned long long pairs = 0;
    while (cur < n_items) {
        const uint32_t nxt = fetch();
        const uint4 itn = nxt < n_items ? a.items[nxt] : make_uint4(0, 0, 0, 0);
        bool nxt_ready = false;
        const uint32_t pool_off = it.x, tcount = it.y, pbeg = it.z, pcnt = it.w;
        if (tcount > 0) {
            if (!cur_ready) issue_chunk(pool_off, min(static_cast<uint32_t>(EV_CHUNK), tcount));
            const float4 P0 = a.sorted[pbeg + min(static_cast<uint32_t>(lane), pcnt - 1)];
            const float4 P1 = a.sorted[pbeg + min(static_cast<uint32_t>(lane) + 32u, pcnt - 1)];
            unsigned long long best0 = KEY_NONE, best1 = KEY_NONE;
            const uint32_t nchunks = (tcount + EV_CHUNK - 1) / EV_CHUNK;
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
#pragma unroll 2
                for (uint32_t j = 0; j < cnt; ++j) {
                    const float4 ca = ws.A[s][j];
                    const float4 cb = ws.B[s][j];
                    const uint32_t ci = static_cast<uint32_t>(ws.I[s][j]);
                    const float d0 = eval_pair<GUARD, NFMA, false>(P0.x, P0.y, P0.z, ca, cb, a.atol, a.eps, nullptr);
                    const float d1 = eval_pair<GUARD, NFMA, false>(P1.x, P1.y, P1.z, ca, cb, a.atol, a.eps, nullptr);
                    const unsigned long long k0 = make_key(d0, ci), k1 = make_key(d1, ci);
                    best0 = k0 < best0 ? k0 : best0;
                    best1 = k1 < best1 ? k1 : best1;
                }
                __syncwarp();            // all lanes are done with stage s before it is armed again
            }
            // cylinders that cannot be pruned (non-finite / non-unit axis): evaluated for every point
            for (uint32_t e = 0; e < a.n_special; ++e) {
                const uint32_t ci = static_cast<uint32_t>(a.special[e]);
                const float4 ca = a.recA[ci], cb = a.recB[ci];
                const float d0 = eval_pair<GUARD, NFMA, false>(P0.x, P0.y, P0.z, ca, cb, a.atol, a.eps, nullptr);
                const float d1 = eval_pair<GUARD, NFMA, false>(P1.x, P1.y, P1.z, ca, cb, a.atol, a.eps, nullptr);
                const unsigned long long k0 = make_key(d0, ci), k1 = make_key(d1, ci);
                best0 = k0 < best0 ? k0 : best0;
                best1 = k1 < best1 ? k1 : best1;
            }
            pairs += static_cast<unsigned long long>(pcnt) * (tcount + a.n_special);
            // ---- fused epilogue: winner-only geometry, label + offset at the original row ----
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t li = static_cast<uint32_t>(lane) + 32u * k;
                if (li < pcnt) {
                    const float4 P = k ? P1 : P0;
                    const uint32_t j = key_index(k ? best1 : best0);
                    const float4 ca = a.recA[j], cb = a.recB[j];
                    PairGeom gm;
                    eval_pair<GUARD, NFMA, true>(P.x, P.y, P.z, ca, cb, a.atol, a.eps, &gm);
                    float ox, oy, oz;
                    mantle_offset<NFMA>(gm, P.x, P.y, P.z, a.move_to_mantle != 0, ox, oy, oz);
                    const int64_t row = static_cast<int64_t>(__float_as_int(P.w));
                    if (a.out_index) a.out_index[row] = static_cast<int32_t>(j);
                    if (a.out_id) a.out_id[row] = a.ids[j];
                    if (a.out_dist) a.out_dist[row] = gm.dist;
                    if (a.out_offset) { a.out_offset[3 * row] = ox; a.out_offset[3 * row + 1] = oy; a.out_offset[3 * row + 2] = oz; }
                    if (a.out_radius) a.out_radius[row] = cb.w;
                }
            }
        }
        cur = nxt;
        it = itn;
        cur_ready = nxt_ready;
    }
    if (lane == 0 && pairs) atomicAdd(&a.st->pairs_grid, pairs);
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
    h->stats.points_grid = static_cast<uint64_t>(a.n);      // minus the outliers, resolved in tm_get_stats
    if (h->n_listed == 0 && h->n_long == 0) return label_brute(h, a);     // only special cylinders: nothing to prune with

    float maxabs = 0.f;
    {
        const GridDesc &d = h->grid;
        maxabs = std::max({std::fabs(d.ox), std::fabs(d.oy), std::fabs(d.oz), std::fabs(d.ox + d.nx * d.h),
                           std::fabs(d.oy + d.ny * d.h), std::fabs(d.oz + d.nz * d.h)});
    }
    const float slack = 32.f * 5.96e-8f * maxabs + 2e-6f;
    const GridDev g = to_dev(h->grid, slack);
    const uint32_t ncodes = h->grid.ncell_codes;
    const size_t n = static_cast<size_t>(a.n);

    // scratch
    const size_t max_occ = std::min<size_t>(n, ncodes);
    const size_t max_items = n / PTS_PER_ITEM + max_occ + 1;
    size_t pool = std::max<size_t>(size_t(1) << 21, std::min<size_t>(4 * n, size_t(1) << 26));
    pool = (pool + 3) & ~size_t(3);
    TM_CUDA(h, h->cell_count.ensure(sizeof(uint32_t) * (static_cast<size_t>(ncodes) + 1)));
    TM_CUDA(h, h->cell_start.ensure(sizeof(uint32_t) * (static_cast<size_t>(ncodes) + 1)));
    TM_CUDA(h, h->pt_cell.ensure(sizeof(uint32_t) * n));
    TM_CUDA(h, h->pt_rank.ensure(sizeof(uint32_t) * n));
    TM_CUDA(h, h->sorted_pts.ensure(sizeof(float4) * n));
    TM_CUDA(h, h->occ_cells.ensure(sizeof(uint32_t) * max_occ));
    TM_CUDA(h, h->tile_meta.ensure(sizeof(uint32_t) * max_occ));
    TM_CUDA(h, h->items.ensure(sizeof(uint4) * max_items));
    TM_CUDA(h, h->outlier_idx.ensure(sizeof(int32_t) * n));
    TM_CUDA(h, h->keys.ensure(sizeof(unsigned long long) * n));
    TM_CUDA(h, h->tileA.ensure(sizeof(float4) * pool));
    TM_CUDA(h, h->tileB.ensure(sizeof(float4) * pool));
    TM_CUDA(h, h->tileI.ensure(sizeof(int32_t) * pool));
    TM_CUDA(h, h->dstats.ensure(sizeof(DevStats) + 64));
    h->tile_pool_entries = pool;
    DevStats *dst = h->dstats.as<DevStats>();
    unsigned int *cursor = reinterpret_cast<unsigned int *>(h->dstats.as<unsigned char>() + sizeof(DevStats) + 16);
    TM_CUDA(h, cudaMemsetAsync(h->dstats.p, 0, sizeof(DevStats) + 64, st));
    TM_CUDA(h, cudaMemsetAsync(h->cell_count.p, 0, sizeof(uint32_t) * ncodes, st));

    const int pt_blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(h->sm_count) * 32));
    bin_count_kernel<<<pt_blocks, 256, 0, st>>>(a.pts, a.n, a.row_stride, g, h->cell_count.as<uint32_t>(),
                                                h->pt_cell.as<uint32_t>(), h->pt_rank.as<uint32_t>(),
                                                h->outlier_idx.as<int32_t>(), h->keys.as<unsigned long long>(), dst);
    TM_CUDA(h, cudaGetLastError());
    mark(h, 1, st);
    int rc = run_scan(h, h->cell_count.as<uint32_t>(), ncodes, 1, h->cell_start.as<uint32_t>(), h->occ_cells.as<uint32_t>(),
                      h->tile_meta.as<uint32_t>(), dst, st);
    if (rc != TM_OK) return rc;
    mark(h, 2, st);
    bin_scatter_kernel<<<pt_blocks, 256, 0, st>>>(a.pts, a.n, a.row_stride, h->pt_cell.as<uint32_t>(),
                                                  h->pt_rank.as<uint32_t>(), h->cell_start.as<uint32_t>(),
                                                  h->sorted_pts.as<float4>());
    TM_CUDA(h, cudaGetLastError());

    mark(h, 3, st);
    TileBuildArgs tb;
    tb.occ_cells = h->occ_cells.as<uint32_t>();
    tb.occ_item_start = h->tile_meta.as<uint32_t>();
    tb.pt_count = h->cell_count.as<uint32_t>();
    tb.pt_start = h->cell_start.as<uint32_t>();
    tb.cyl_start = h->cyl_cell_start.as<uint32_t>();
    tb.cyl_list = h->cyl_cell_list.as<int32_t>();
    tb.long_list = h->long_list.as<int32_t>();
    tb.n_long = h->n_long;
    tb.recA = h->recA.as<float4>(); tb.recB = h->recB.as<float4>();
    tb.boxlo = h->boxlo.as<float4>(); tb.boxhi = h->boxhi.as<float4>();
    tb.sorted = h->sorted_pts.as<float4>();
    tb.tileA = h->tileA.as<float4>(); tb.tileB = h->tileB.as<float4>(); tb.tileI = h->tileI.as<int32_t>();
    tb.pool_entries = static_cast<uint32_t>(pool);
    tb.items = h->items.as<uint4>();
    tb.outlier_idx = h->outlier_idx.as<int32_t>();
    tb.outlier_keys = h->keys.as<unsigned long long>();
    tb.st = dst;
    const int tb_blocks = static_cast<int>(std::min<size_t>((max_occ + TB_WARPS - 1) / TB_WARPS, static_cast<size_t>(h->sm_count) * 8));
    tile_build_kernel<<<tb_blocks, TB_WARPS * 32, 0, st>>>(tb, g);
    TM_CUDA(h, cudaGetLastError());

    mark(h, 4, st);
    EvalArgs ev;
    ev.items = h->items.as<uint4>();
    ev.cursor = cursor;
    ev.sorted = h->sorted_pts.as<float4>();
    ev.tileA = h->tileA.as<float4>(); ev.tileB = h->tileB.as<float4>(); ev.tileI = h->tileI.as<int32_t>();
    ev.recA = h->recA.as<float4>(); ev.recB = h->recB.as<float4>();
    ev.ids = h->ids.as<int32_t>();
    ev.special = h->special.as<int32_t>();
    ev.n_special = h->n_special;
    ev.atol = a.prm.perp_atol; ev.eps = a.prm.norm_eps;
    ev.move_to_mantle = a.prm.move_to_mantle;
    ev.out_index = a.out_index; ev.out_id = a.out_id; ev.out_dist = a.out_dist; ev.out_offset = a.out_offset;
    ev.out_radius = a.out_radius;
    ev.st = dst;
    const size_t ev_smem = sizeof(WarpStage) * EV_WARPS + sizeof(uint64_t) * 2 * EV_WARPS;
    const bool guard = a.prm.norm_eps > 0.f, nfma = a.prm.norm_fma != 0;
    const int ev_blocks = h->sm_count * 3;
#define TM_EVAL_CASE(G, F)                                                                                         \
    do {                                                                                                           \
        TM_CUDA(h, cudaFuncSetAttribute(evaluate_kernel<G, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                        static_cast<int>(ev_smem)));                                               \
        evaluate_kernel<G, F><<<ev_blocks, EV_WARPS * 32, ev_smem, st>>>(ev);                                      \
    } while (0)
    if (guard) { if (nfma) TM_EVAL_CASE(true, true); else TM_EVAL_CASE(true, false); }
    else       { if (nfma) TM_EVAL_CASE(false, true); else TM_EVAL_CASE(false, false); }
#undef TM_EVAL_CASE
    TM_CUDA(h, cudaGetLastError());

    // everything the grid could not answer: exhaustive search, same arithmetic, same outputs
    mark(h, 5, st);
    rc = label_brute_subset(h, a, h->outlier_idx.as<int32_t>(), &dst->outliers, static_cast<unsigned int>(n));
    return rc;
}

}  // namespace tmn

// Point-neighbourhood features that the label / projection drivers append right after the nearest-cylinder path
// (Modules/Features.py:178-229, called at PreProcessing/LabelGenerationCuda.py:197-198 and Modules/Projection.py:420-427):
//
//   compute_normals_ckdtree   (Features.py:111-133)  k = 15 nearest neighbours (the point itself included), covariance of
//                             the neighbours relative to the point (np.cov: mean-subtracted, / (k-1)), SVD
//   compute_curvature_ckdtree (Features.py:136-157)  k = 10, eigenvalues of the same covariance
//   compute_density_ckdtree   (Features.py:160-172)  number of points within 0.1 m (the point itself included)
//
// The reference builds a scipy cKDTree and then loops over the points in Python (np.cov + LAPACK per point): minutes per
// tree.  Here the neighbour search and the covariance run on the device in float64 (the labelled cloud is float64) over a
// uniform grid of the cloud itself; the 3x3 decompositions stay with LAPACK on the host, batched, because the reference's
// output is LAPACK's sign choice (see Modules/Features.py of this package).
//
//   grid:    cell edge from the bounding box and the point count (a few points per occupied cell), counting sort of the
//            points by cell (same RED / scan / atomic-cursor scheme as the cylinder path);
//   search:  one thread per point, Chebyshev rings of cells around the point's cell, a sorted top-k list in local memory;
//            after ring r every unvisited point is farther than r*h + (distance of the point to its own cell's faces), so
//            the search stops as soon as the k-th distance is below that bound: exact k nearest neighbours;
//   output:  covariance (9 doubles) at the point's original row, optionally the neighbour rows.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "tm_core.cuh"

namespace tmn {

struct KnnGrid {
    double ox, oy, oz, h, inv_h;
    int nx, ny, nz;
};

__device__ __forceinline__ long long ordered_of(double v) {
    const long long b = __double_as_longlong(v);
    return b >= 0 ? b : b ^ 0x7fffffffffffffffLL;
}
static inline double ordered_to_double(long long k) {
    const long long b = k >= 0 ? k : k ^ 0x7fffffffffffffffLL;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

// box[0..2] = min, box[3..5] = max (ordered 64-bit keys); box[6] = number of non-finite rows
__global__ void __launch_bounds__(256) knn_bbox_kernel(const double *__restrict__ pts, int64_t n, int64_t row_stride,
                                                       long long *__restrict__ box) {
    long long lo[3] = {0x7fffffffffffffffLL, 0x7fffffffffffffffLL, 0x7fffffffffffffffLL};
    long long hi[3] = {static_cast<long long>(0x8000000000000000ULL), static_cast<long long>(0x8000000000000000ULL),
                       static_cast<long long>(0x8000000000000000ULL)};
    long long bad = 0;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const double *p = pts + i * row_stride;
        if (!(isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]))) { ++bad; continue; }
#pragma unroll
        for (int k = 0; k < 3; ++k) { const long long o = ordered_of(p[k]); lo[k] = min(lo[k], o); hi[k] = max(hi[k], o); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = min(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = max(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(&box[k], lo[k]); atomicMax(&box[3 + k], hi[k]); }
        if (bad) atomicAdd(reinterpret_cast<unsigned long long *>(&box[6]), static_cast<unsigned long long>(bad));
    }
}

__device__ __forceinline__ int knn_coord(double p, double o, double inv_h, int n) {
    const int c = static_cast<int>(floor((p - o) * inv_h));
    return min(max(c, 0), n - 1);
}
__device__ __forceinline__ uint32_t knn_cell(const KnnGrid &g, double x, double y, double z) {
    const int cx = knn_coord(x, g.ox, g.inv_h, g.nx), cy = knn_coord(y, g.oy, g.inv_h, g.ny), cz = knn_coord(z, g.oz, g.inv_h, g.nz);
    return (static_cast<uint32_t>(cz) * g.ny + cy) * g.nx + cx;
}

__global__ void __launch_bounds__(256) knn_count_kernel(const double *__restrict__ pts, int64_t n, int64_t row_stride, KnnGrid g,
                                                        uint32_t *__restrict__ cells) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const double *p = pts + i * row_stride;
        atomicAdd(&cells[knn_cell(g, p[0], p[1], p[2])], 1u);
    }
}

// cells[] is zeroed again before this pass and serves as the per-cell cursor
__global__ void __launch_bounds__(256) knn_scatter_kernel(const double *__restrict__ pts, int64_t n, int64_t row_stride, KnnGrid g,
                                                          uint32_t *__restrict__ cells, const uint32_t *__restrict__ start,
                                                          double4 *__restrict__ sorted) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const double *p = pts + i * row_stride;
        const double x = p[0], y = p[1], z = p[2];
        const uint32_t c = knn_cell(g, x, y, z);
        const uint32_t pos = start[c] + atomicAdd(&cells[c], 1u);
        sorted[pos] = make_double4(x, y, z, __longlong_as_double(i));
    }
}

constexpr int KNN_MAX = 32;

// exact k nearest neighbours + covariance, one thread per point (in cell order: neighbouring threads walk the same cells)
__global__ void __launch_bounds__(128) knn_cov_kernel(const double4 *__restrict__ sorted, const uint32_t *__restrict__ start, int64_t n,
                                                      KnnGrid g, int k, double *__restrict__ out_cov, int32_t *__restrict__ out_idx) {
    const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (t >= n) return;
    const double4 q = sorted[t];
    const int64_t row = __double_as_longlong(q.w);
    const int cx = knn_coord(q.x, g.ox, g.inv_h, g.nx), cy = knn_coord(q.y, g.oy, g.inv_h, g.ny), cz = knn_coord(q.z, g.oz, g.inv_h, g.nz);
    double bd[KNN_MAX];
    uint32_t bi[KNN_MAX];        // position in `sorted`
    int have = 0;
    // distance from the point to the faces of its own cell: what a ring adds to r*h
    const double fx = q.x - (g.ox + cx * g.h), fy = q.y - (g.oy + cy * g.h), fz = q.z - (g.oz + cz * g.h);
    const double margin = fmax(0.0, fmin(fmin(fmin(fx, g.h - fx), fmin(fy, g.h - fy)), fmin(fz, g.h - fz)));
    const int rmax = max(max(g.nx, g.ny), g.nz);
    // candidates of the cells [xa, xb] of row (y, z): one contiguous run of the sorted array
    auto visit = [&](int xa, int xb, int y, int z) {
        const uint32_t row0 = (static_cast<uint32_t>(z) * g.ny + y) * g.nx;
        for (uint32_t j = start[row0 + xa], e = start[row0 + xb + 1]; j < e; ++j) {
            const double4 p = sorted[j];
            const double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (have == k && !(d2 < bd[k - 1] || (d2 == bd[k - 1] && __double_as_longlong(p.w) < __double_as_longlong(sorted[bi[k - 1]].w)))) continue;
            // insert into the ascending list (ties: lower original row first)
            int pos = have < k ? have : k - 1;
            while (pos > 0 && (bd[pos - 1] > d2 || (bd[pos - 1] == d2 && __double_as_longlong(sorted[bi[pos - 1]].w) > __double_as_longlong(p.w)))) {
                bd[pos] = bd[pos - 1];
                bi[pos] = bi[pos - 1];
                --pos;
            }
            bd[pos] = d2;
            bi[pos] = j;
            if (have < k) ++have;
        }
    };
    for (int r = 0; r <= rmax; ++r) {
        const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nz - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, g.ny - 1);
        const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
        for (int z = z0; z <= z1; ++z) {
            for (int y = y0; y <= y1; ++y) {
                if (abs(z - cz) == r || abs(y - cy) == r) {
                    visit(x0, x1, y, z);                                     // a face of the shell: the whole row
                } else {                                                     // interior row: only its two end cells
                    if (cx - r >= 0) visit(cx - r, cx - r, y, z);
                    if (cx + r < g.nx) visit(cx + r, cx + r, y, z);
                }
            }
        }
        if (have == k) {
            const double bound = r * g.h + margin;
            if (bd[k - 1] <= bound * bound) break;
        }
        if (z0 == 0 && y0 == 0 && x0 == 0 && z1 == g.nz - 1 && y1 == g.ny - 1 && x1 == g.nx - 1) break;     // whole grid visited
    }
    // np.cov of (neighbours - point): rows = coordinates, mean-subtracted, / (k - 1)   (Features.py:127-128)
    double mx = 0, my = 0, mz = 0;
    for (int i = 0; i < have; ++i) {
        const double4 p = sorted[bi[i]];
        mx += p.x - q.x; my += p.y - q.y; mz += p.z - q.z;
    }
    mx /= have; my /= have; mz /= have;
    double cxx = 0, cxy = 0, cxz = 0, cyy = 0, cyz = 0, czz = 0;
    for (int i = 0; i < have; ++i) {
        const double4 p = sorted[bi[i]];
        const double ax = (p.x - q.x) - mx, ay = (p.y - q.y) - my, az = (p.z - q.z) - mz;
        cxx += ax * ax; cxy += ax * ay; cxz += ax * az; cyy += ay * ay; cyz += ay * az; czz += az * az;
    }
    const double f = 1.0 / (have - 1);
    double *o = out_cov + 9 * row;
    o[0] = cxx * f; o[1] = cxy * f; o[2] = cxz * f;
    o[3] = cxy * f; o[4] = cyy * f; o[5] = cyz * f;
    o[6] = cxz * f; o[7] = cyz * f; o[8] = czz * f;
    if (out_idx)
        for (int i = 0; i < k; ++i) out_idx[row * k + i] = i < have ? static_cast<int32_t>(__double_as_longlong(sorted[bi[i]].w)) : -1;
}

// number of points within `radius` of each point (itself included), cell edge >= radius so 27 cells suffice
__global__ void __launch_bounds__(128) radius_count_kernel(const double4 *__restrict__ sorted, const uint32_t *__restrict__ start, int64_t n,
                                                           KnnGrid g, double radius, int32_t *__restrict__ out_count) {
    const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (t >= n) return;
    const double4 q = sorted[t];
    const int cx = knn_coord(q.x, g.ox, g.inv_h, g.nx), cy = knn_coord(q.y, g.oy, g.inv_h, g.ny), cz = knn_coord(q.z, g.oz, g.inv_h, g.nz);
    const double r2 = radius * radius;
    int cnt = 0;
    for (int z = max(cz - 1, 0); z <= min(cz + 1, g.nz - 1); ++z)
        for (int y = max(cy - 1, 0); y <= min(cy + 1, g.ny - 1); ++y) {
            const uint32_t c0 = (static_cast<uint32_t>(z) * g.ny + y) * g.nx + max(cx - 1, 0);
            const uint32_t c1 = (static_cast<uint32_t>(z) * g.ny + y) * g.nx + min(cx + 1, g.nx - 1);
            for (uint32_t j = start[c0], e = start[c1 + 1]; j < e; ++j) {       // the three x-cells are one contiguous run
                const double4 p = sorted[j];
                const double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
                cnt += (dx * dx + dy * dy + dz * dz <= r2) ? 1 : 0;
            }
        }
    out_count[__double_as_longlong(q.w)] = cnt;
}

__global__ void __launch_bounds__(256) knn_occupied_kernel(const uint32_t *__restrict__ cells, uint32_t ncells, unsigned int *__restrict__ occ) {
    unsigned int mine = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ncells; i += gridDim.x * blockDim.x) mine += cells[i] ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(occ, mine);
}

// bounding box -> grid -> counting sort.  min_cell: lower bound on the cell edge (radius search), 0 = automatic.
static int knn_build(tm_handle *h, const double *pts, int64_t n, int64_t row_stride, double min_cell, double pts_per_cell, KnnGrid *out,
                     cudaStream_t st) {
    TM_CUDA(h, h->knn_box.ensure(sizeof(long long) * 8));
    const long long init[8] = {0x7fffffffffffffffLL, 0x7fffffffffffffffLL, 0x7fffffffffffffffLL,
                               static_cast<long long>(0x8000000000000000ULL), static_cast<long long>(0x8000000000000000ULL),
                               static_cast<long long>(0x8000000000000000ULL), 0, 0};
    TM_CUDA(h, cudaMemcpyAsync(h->knn_box.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(h->sm_count) * 16));
    knn_bbox_kernel<<<blocks, 256, 0, st>>>(pts, n, row_stride, h->knn_box.as<long long>());
    TM_KCHECK(h, st, "knn_bbox_kernel");
    long long box[8];
    TM_CUDA(h, cudaMemcpyAsync(box, h->knn_box.p, sizeof(box), cudaMemcpyDeviceToHost, st));
    TM_CUDA(h, cudaStreamSynchronize(st));
    if (box[6] != 0) return fail(h, TM_ERR_INVALID, "point features: the cloud holds non-finite coordinates%s%s");
    double lo[3], hi[3];
    for (int k = 0; k < 3; ++k) { lo[k] = ordered_to_double(box[k]); hi[k] = ordered_to_double(box[3 + k]); }
    // cell edge: pts_per_cell points per cell if the cloud filled its box; surface-like clouds fill a few per cent of it,
    // so occupied cells end up with a handful of points each
    double ext[3], vol = 1.0;
    for (int k = 0; k < 3; ++k) { ext[k] = std::max(hi[k] - lo[k], 1e-9); vol *= ext[k]; }
    if (const char *env = getenv("TM_KNN_BOX_PER_CELL")) { const double v = atof(env); if (v >= 0.01 && v <= 256.0) pts_per_cell = v; }
    double hcell = std::cbrt(vol * pts_per_cell / static_cast<double>(n));
    const double longest = std::max(ext[0], std::max(ext[1], ext[2]));
    hcell = std::max(hcell, longest / 2048.0);
    hcell = std::max(hcell, min_cell);
    KnnGrid g;
    uint32_t ncells = 0;
    for (int pass = 0; pass < 2; ++pass) {
        for (;;) {
            g.nx = static_cast<int>(ext[0] / hcell) + 1; g.ny = static_cast<int>(ext[1] / hcell) + 1; g.nz = static_cast<int>(ext[2] / hcell) + 1;
            if (static_cast<double>(g.nx) * g.ny * g.nz <= static_cast<double>(1u << 27)) break;
            hcell *= 1.26;
        }
        g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2]; g.h = hcell; g.inv_h = 1.0 / hcell;
        ncells = static_cast<uint32_t>(g.nx) * g.ny * g.nz;
        TM_CUDA(h, h->knn_cells.ensure(sizeof(uint32_t) * (static_cast<size_t>(ncells) + 1)));
        TM_CUDA(h, cudaMemsetAsync(h->knn_cells.p, 0, sizeof(uint32_t) * ncells, st));
        knn_count_kernel<<<blocks, 256, 0, st>>>(pts, n, row_stride, g, h->knn_cells.as<uint32_t>());
        TM_KCHECK(h, st, "knn_count_kernel");
        if (pass == 1 || min_cell > 0.0) break;
        // a surface-sampled cloud fills a few per cent of its box: measure the points per OCCUPIED cell and shrink the
        // edge (points per occupied cell of a surface go with h^2) until there are a handful
        unsigned int occ = 0;
        TM_CUDA(h, cudaMemsetAsync(h->knn_box.as<long long>() + 7, 0, sizeof(long long), st));
        knn_occupied_kernel<<<h->sm_count * 8, 256, 0, st>>>(h->knn_cells.as<uint32_t>(), ncells,
                                                             reinterpret_cast<unsigned int *>(h->knn_box.as<long long>() + 7));
        TM_CUDA(h, cudaMemcpyAsync(&occ, h->knn_box.as<long long>() + 7, sizeof(occ), cudaMemcpyDeviceToHost, st));
        TM_CUDA(h, cudaStreamSynchronize(st));
        const double per_occ = static_cast<double>(n) / std::max(1u, occ);
        double target = 12.0;
        if (const char *env = getenv("TM_KNN_PER_CELL")) { const double v = atof(env); if (v >= 0.5 && v <= 256.0) target = v; }
        if (per_occ <= 2.0 * target) break;
        hcell = std::max(hcell * std::sqrt(target / per_occ), longest / 2048.0);
    }
    TM_CUDA(h, h->knn_start.ensure(sizeof(uint32_t) * (static_cast<size_t>(ncells) + 1)));
    TM_CUDA(h, h->knn_sorted.ensure(sizeof(double4) * static_cast<size_t>(n)));
    int rc = exclusive_scan_u32(h, h->knn_cells.as<uint32_t>(), ncells, h->knn_start.as<uint32_t>(), st);
    if (rc != TM_OK) return rc;
    TM_CUDA(h, cudaMemsetAsync(h->knn_cells.p, 0, sizeof(uint32_t) * ncells, st));
    knn_scatter_kernel<<<blocks, 256, 0, st>>>(pts, n, row_stride, g, h->knn_cells.as<uint32_t>(), h->knn_start.as<uint32_t>(),
                                               h->knn_sorted.as<double4>());
    TM_KCHECK(h, st, "knn_scatter_kernel");
    *out = g;
    return TM_OK;
}

}  // namespace tmn

using namespace tmn;

extern "C" {

int tm_knn_covariance(tm_handle *h, const double *pts, int64_t n, int64_t row_stride, int32_t k, double *out_cov, int32_t *out_idx,
                      void *stream) {
    if (!h) return TM_ERR_INVALID;
    if (n < 0 || k < 2 || k > KNN_MAX) return fail(h, TM_ERR_INVALID, "tm_knn_covariance: k must be in [2, 32] and n >= 0%s%s");
    if (n == 0) return TM_OK;
    if (!pts || !out_cov || row_stride < 3) return fail(h, TM_ERR_INVALID, "tm_knn_covariance: null pointer or row_stride < 3%s%s");
    if (n < k) return fail(h, TM_ERR_INVALID, "tm_knn_covariance: fewer points than neighbours requested%s%s");
    if (n > 0x7fffffffLL) return fail(h, TM_ERR_INVALID, "tm_knn_covariance: more than 2^31-1 points%s%s");
    TM_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    KnnGrid g;
    int rc = knn_build(h, pts, n, row_stride, 0.0, 1.0, &g, st);
    if (rc != TM_OK) return rc;
    knn_cov_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(h->knn_sorted.as<double4>(), h->knn_start.as<uint32_t>(), n, g, k,
                                                                         out_cov, out_idx);
    TM_KCHECK(h, st, "knn_cov_kernel");
    return TM_OK;
}

int tm_radius_count(tm_handle *h, const double *pts, int64_t n, int64_t row_stride, double radius, int32_t *out_count, void *stream) {
    if (!h) return TM_ERR_INVALID;
    if (n < 0 || !(radius > 0.0)) return fail(h, TM_ERR_INVALID, "tm_radius_count: radius must be > 0 and n >= 0%s%s");
    if (n == 0) return TM_OK;
    if (!pts || !out_count || row_stride < 3) return fail(h, TM_ERR_INVALID, "tm_radius_count: null pointer or row_stride < 3%s%s");
    if (n > 0x7fffffffLL) return fail(h, TM_ERR_INVALID, "tm_radius_count: more than 2^31-1 points%s%s");
    TM_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    KnnGrid g;
    int rc = knn_build(h, pts, n, row_stride, radius * (1.0 + 1e-9), 1.0, &g, st);
    if (rc != TM_OK) return rc;
    radius_count_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(h->knn_sorted.as<double4>(), h->knn_start.as<uint32_t>(), n, g,
                                                                              radius, out_count);
    TM_KCHECK(h, st, "radius_count_kernel");
    return TM_OK;
}

}  // extern "C"

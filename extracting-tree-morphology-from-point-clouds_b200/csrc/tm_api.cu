// C ABI of the library (include/treemorph_nn.h): handle management, cylinder preparation and
// packing, dispatch between the exhaustive and voxel-grid kernels, record assembly and the
// pipelined host entry point.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <chrono>
#include <thread>
#include <vector>

#include <sched.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "tm_core.cuh"
#include "tm_eval.cuh"

namespace tmn {

// ------------------------------------------------------------------------------------------------
// cylinder preparation: LabelGenerationCuda.py:121-123 / Projection.py:126-132
// ------------------------------------------------------------------------------------------------
template <bool NFMA>
__global__ void prepare_kernel(const float *__restrict__ start, int64_t s_rs, int64_t s_cs, const float *__restrict__ end,
                               int64_t e_rs, int64_t e_cs, int64_t m, float axis_eps, float *__restrict__ out_len,
                               float *__restrict__ out_unit) {
    const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (c >= m) return;
    // axis = end - start
    const float ax = sub(end[c * e_rs], start[c * s_rs]);
    const float ay = sub(end[c * e_rs + e_cs], start[c * s_rs + s_cs]);
    const float az = sub(end[c * e_rs + 2 * e_cs], start[c * s_rs + 2 * s_cs]);
    // axis_length = torch.norm(axis, dim=1, keepdim=True)
    const float len = norm3<NFMA>(ax, ay, az);
    // Projection.py:129-131: safe_axis_length[safe_axis_length < eps] = eps   (variant A: no guard)
    const float dv = (axis_eps > 0.f && len < axis_eps) ? axis_eps : len;
    out_len[c] = len;                       // the UNguarded length is what the kernel receives (Projection.py:137)
    out_unit[3 * c] = __fdiv_rn(ax, dv);
    out_unit[3 * c + 1] = __fdiv_rn(ay, dv);
    out_unit[3 * c + 2] = __fdiv_rn(az, dv);
}

// ------------------------------------------------------------------------------------------------
// cylinder table packing
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}

// One thread per cylinder: two float4 records, the solid-cylinder AABB, and the classification
//   regular : finite, unit (or zero) axis  → participates in the voxel index
//   special : anything else (NaN axis of a zero-length cylinder in variant A, ...) → cannot be
//             pruned, evaluated for every point
__global__ void pack_kernel(const float *__restrict__ start, int64_t s_rs, int64_t s_cs, const float *__restrict__ unit,
                            int64_t u_rs, int64_t u_cs, const float *__restrict__ length, int64_t l_s,
                            const float *__restrict__ radius, int64_t r_s, const int32_t *__restrict__ ids, int64_t i_s,
                            int m, float4 *__restrict__ recA, float4 *__restrict__ recB, float4 *__restrict__ recAB,
                            int32_t *__restrict__ out_ids,
                            float4 *__restrict__ boxlo, float4 *__restrict__ boxhi, int *__restrict__ bbox,
                            int32_t *__restrict__ special, int32_t *__restrict__ aligned) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const float sx = start[c * s_rs], sy = start[c * s_rs + s_cs], sz = start[c * s_rs + 2 * s_cs];
    const float ux = unit[c * u_rs], uy = unit[c * u_rs + u_cs], uz = unit[c * u_rs + 2 * u_cs];
    const float len = length[c * l_s], rad = radius[c * r_s];
    recA[c] = make_float4(sx, sy, sz, len);
    recB[c] = make_float4(ux, uy, uz, rad);
    recAB[2 * c] = make_float4(sx, sy, sz, len);
    recAB[2 * c + 1] = make_float4(ux, uy, uz, rad);
    out_ids[c] = ids ? ids[c * i_s] : c;
    const bool finite = isfinite(sx) && isfinite(sy) && isfinite(sz) && isfinite(ux) && isfinite(uy) && isfinite(uz) &&
                        isfinite(len) && isfinite(rad);
    const float un = ux * ux + uy * uy + uz * uz;
    // the capsule bound of tm_eval.cuh needs a unit axis: a correctly normalised fp32 vector is within 4e-7
    const bool unit_ok = fabsf(un - 1.f) <= 4e-6f || (un == 0.f && len == 0.f);
    if (!finite || !unit_ok || len < 0.f) {
        boxlo[c] = make_float4(0.f, 0.f, 0.f, 1.f);
        boxhi[c] = make_float4(0.f, 0.f, 0.f, 1.f);
        const int s = atomicAdd(&bbox[6], 1);
        special[s] = c;
        return;
    }
    const float ex = fmaf(len, ux, sx), ey = fmaf(len, uy, sy), ez = fmaf(len, uz, sz);
    const float ar = fabsf(rad);
    float lx = fminf(sx, ex) - ar, ly = fminf(sy, ey) - ar, lz = fminf(sz, ez) - ar;
    float hx = fmaxf(sx, ex) + ar, hy = fmaxf(sy, ey) + ar, hz = fmaxf(sz, ez) + ar;
    const float px = 1e-6f * fmaxf(fabsf(lx), fabsf(hx)) + 1e-6f;
    const float py = 1e-6f * fmaxf(fabsf(ly), fabsf(hy)) + 1e-6f;
    const float pz = 1e-6f * fmaxf(fabsf(lz), fabsf(hz)) + 1e-6f;
    lx -= px; ly -= py; lz -= pz; hx += px; hy += py; hz += pz;
    boxlo[c] = make_float4(lx, ly, lz, 0.f);
    boxhi[c] = make_float4(hx, hy, hz, 0.f);
    atomicMin(&bbox[0], float_to_ordered(lx)); atomicMin(&bbox[1], float_to_ordered(ly)); atomicMin(&bbox[2], float_to_ordered(lz));
    atomicMax(&bbox[3], float_to_ordered(hx)); atomicMax(&bbox[4], float_to_ordered(hy)); atomicMax(&bbox[5], float_to_ordered(hz));
    atomicAdd(&bbox[7], 1);
    // size statistics for the automatic voxel edge: sums of the lengths and of the diameters in 2^-20 m units (integers:
    // order independent, so the edge does not change from run to run)
    atomicAdd(reinterpret_cast<unsigned long long *>(bbox + 10), static_cast<unsigned long long>(fminf(len, 1.0e6f) * 1048576.0f));
    atomicAdd(reinterpret_cast<unsigned long long *>(bbox + 12), static_cast<unsigned long long>(fminf(2.f * ar, 1.0e6f) * 1048576.0f));
    // axis-parallel (two exactly-zero unit components): NaN for every point on the axis line in variant A
    if ((ux == 0.f) + (uy == 0.f) + (uz == 0.f) >= 2) aligned[atomicAdd(&bbox[8], 1)] = c;
}

// ------------------------------------------------------------------------------------------------
// (N,7) float64 record of generate_offset_cloud_cuda_batched (LabelGenerationCuda.py:114,131-133)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void assemble_kernel(const T *__restrict__ cloud, int64_t n, int64_t row_stride, const float *__restrict__ offset,
                                const int32_t *__restrict__ id, double *__restrict__ out) {
    // 7 consecutive threads write one row → fully coalesced 8-byte stores
    const int64_t total = n * 7;
    for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = e / 7;
        const int col = static_cast<int>(e - row * 7);
        double v;
        if (col < 3) v = static_cast<double>(cloud[row * row_stride + col]);       // xyz at the caller's precision
        else if (col < 6) v = static_cast<double>(offset[3 * row + (col - 3)]);
        else v = static_cast<double>(id[row]);
        out[e] = v;
    }
}

__global__ void f64_to_f32_xyz_kernel(const double *__restrict__ cloud, int64_t n, int64_t row_stride, float *__restrict__ out) {
    const int64_t total = n * 3;
    for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = e / 3;
        const int col = static_cast<int>(e - row * 3);
        out[e] = static_cast<float>(cloud[row * row_stride + col]);               // torch.tensor(points, dtype=float32): RNE
    }
}

__global__ void f32_xyz_kernel(const float *__restrict__ cloud, int64_t n, int64_t row_stride, float *__restrict__ out) {
    const int64_t total = n * 3;
    for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = e / 3;
        out[e] = cloud[row * row_stride + (e - row * 3)];
    }
}

// ------------------------------------------------------------------------------------------------
// FP32 pipe probe
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) fp32_probe_kernel(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
        if (KIND == 0) {            // FFMA, register operands
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        } else {                    // alternating FMUL / FADD, the instruction mix of the distance kernel
            x0 = __fmul_rn(x0, a); x1 = __fadd_rn(x1, b); x2 = __fmul_rn(x2, a); x3 = __fadd_rn(x3, b);
            x4 = __fmul_rn(x4, a); x5 = __fadd_rn(x5, b); x6 = __fmul_rn(x6, a); x7 = __fadd_rn(x7, b);
        }
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;   // never true in practice; keeps the chains alive
}

// ------------------------------------------------------------------------------------------------
// arithmetic self-test: div3 / sqrt_rn against the compiler's IEEE routines
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// operand classes: 0 = the kernel's regime (components of a vector over its norm), 1 = random bit patterns
// (every exponent, denormals, NaN, inf), 2 = edge values
__global__ void selftest_kernel(uint64_t n, uint32_t seed, unsigned long long *__restrict__ bad) {
    unsigned long long local = 0;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t a = mix32(static_cast<uint32_t>(i) * 4u + seed), b = mix32(a + 1u), c = mix32(b + 2u), d = mix32(c + 3u);
        float x, y, z, rho;
        const uint32_t cls = static_cast<uint32_t>(i % 3u);
        // numerators stay in div3's documented domain: |numerator| <= ~|rho| and either 0 or >= 2^-103
        const float f1 = __uint_as_float((b & 0x007fffffu) | 0x3f800000u) - 1.5f;     // [-0.5, 0.5)
        const float f2 = (__uint_as_float((c & 0x007fffffu) | 0x3f800000u) - 1.5f) * 2.f;
        const float f3 = __uint_as_float((d & 0x007fffffu) | 0x3f800000u) - 1.f;      // [0, 1)
        if (cls == 0) {                 // a vector over its own norm, scales 2^-30 .. 2^29
            const float sc = __uint_as_float(((a >> 9) % 60u + 97u) << 23);
            x = f1 * sc; y = f2 * sc; z = f3 * sc;
            rho = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
        } else if (cls == 1) {          // any denominator bit pattern: every exponent, denormals, NaN, inf, negative
            rho = __uint_as_float(a);
            x = __fmul_rn(rho, f1); y = __fmul_rn(rho, f2); z = __fmul_rn(rho, f3);
        } else {                        // edge denominators
            const float edge[8] = {0.f, -0.f, 1e-8f, 1.17549435e-38f, 1e-45f, __int_as_float(0x7f800000), 3.4e38f, 1.f};
            rho = edge[a & 7u];
            x = __fmul_rn(rho, f1); y = (a & 8u) ? rho : 0.f; z = (a & 16u) ? __fmul_rn(rho, f3) : -rho;
        }
        if (fabsf(x) < 9.86e-32f) x = 0.f;
        if (fabsf(y) < 9.86e-32f) y = 0.f;
        if (fabsf(z) < 9.86e-32f) z = 0.f;
        float qx, qy, qz;
        div3(x, y, z, rho, qx, qy, qz);
        const float ex = __fdiv_rn(x, rho), ey = __fdiv_rn(y, rho), ez = __fdiv_rn(z, rho);
        auto same = [](float p, float q) { return __float_as_uint(p) == __float_as_uint(q) || (p != p && q != q); };
        if (!same(qx, ex) || !same(qy, ey) || !same(qz, ez)) ++local;
        // square root: arbitrary bit patterns, IEEE for every operand
        const float sq_in[6] = {fabsf(x), rho, __fmul_rn(x, x), __uint_as_float(b), __uint_as_float(c & 0x7fffffffu), -fabsf(y)};
        for (int k = 0; k < 6; ++k)
            if (!same(sqrt_rn(sq_in[k]), __fsqrt_rn(sq_in[k]))) ++local;
    }
    if (local) atomicAdd(bad, local);
}

}  // namespace tmn

// Persistent host workers (created on first use, joined in tm_destroy).  run(fn) executes fn(worker, workers) on
// every worker including the calling thread and returns when all are done.
struct tmn::HostPool {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    std::function<void(unsigned, unsigned)> job;
    uint64_t generation = 0;
    unsigned pending = 0, n = 1;
    bool stop = false;
    explicit HostPool(unsigned nthreads) : n(std::max(1u, nthreads)) {
        for (unsigned t = 1; t < n; ++t)
            workers.emplace_back([this, t] {
                uint64_t seen = 0;
                for (;;) {
                    std::function<void(unsigned, unsigned)> fn;
                    {
                        std::unique_lock<std::mutex> lk(m);
                        cv_job.wait(lk, [&] { return stop || generation != seen; });
                        if (stop) return;
                        seen = generation;
                        fn = job;
                    }
                    fn(t, n);
                    {
                        std::lock_guard<std::mutex> lk(m);
                        if (--pending == 0) cv_done.notify_one();
                    }
                }
            });
    }
    void run(const std::function<void(unsigned, unsigned)> &fn) {
        {
            std::lock_guard<std::mutex> lk(m);
            job = fn;
            pending = n - 1;
            ++generation;
        }
        cv_job.notify_all();
        fn(0, n);
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv_job.notify_all();
        for (auto &t : workers) t.join();
    }
};

using namespace tmn;

extern "C" {

int tm_version(void) { return TM_ABI_VERSION; }

const char *tm_status_string(int status) {
    switch (status) {
        case TM_OK: return "ok";
        case TM_ERR_INVALID: return "invalid argument";
        case TM_ERR_NO_CYLINDERS: return "argmin(): expected reduction dim 1 to have non-zero size (no cylinders)";
        case TM_ERR_CUDA: return "CUDA error";
        case TM_ERR_NOMEM: return "out of memory";
        case TM_ERR_STATE: return "call order violated";
        default: return "unknown status";
    }
}

const char *tm_last_error(const tm_handle *h) { return h ? h->err : "null handle"; }

int tm_create(int device, tm_handle **out) {
    if (!out) return TM_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return TM_ERR_CUDA;
    tm_handle *h = new (std::nothrow) tm_handle();
    if (!h) return TM_ERR_NOMEM;
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return TM_ERR_CUDA; }
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) h->sm_count = sms;
    *out = h;
    return TM_OK;
}

int tm_destroy(tm_handle *h) {
    if (!h) return TM_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    tmn::DevBuf *bufs[] = {&h->recA, &h->recB, &h->recAB, &h->cells, &h->bvh_nodes, &h->bvh_rows, &h->bvh_leafAB, &h->knn_cells, &h->knn_start, &h->knn_sorted, &h->knn_box, &h->ids, &h->boxlo, &h->boxhi, &h->bbox, &h->cyl_cell_start, &h->cyl_cell_cnt,
                          &h->cyl_cell_near, &h->tileLB, &h->tile_keys, &h->long_list, &h->special, &h->aligned, &h->keys, &h->cell_count,
                          &h->cell_start, &h->block_sums, &h->sorted_pts, &h->tileAB, &h->tileI, &h->items, &h->warp_item, &h->undecided, &h->late_rows, &h->tile_desc,
                          &h->pend_idx, &h->brute_slots, &h->win, &h->dstats, &h->scratch_f, &h->cloud_res, &h->small_in,
                          &h->small_out, &h->bvh_scratch};
    tm_comm_destroy(h);
    h->comm_table.release();
    h->comm_count.release();
    for (auto *b : bufs) b->release();
    for (int i = 0; i < tmn::PIPE_SLOTS; ++i) {
        h->chunk_in[i].release(); h->chunk_rec[i].release(); h->chunk_off[i].release(); h->chunk_id[i].release();
        h->chunk_dist[i].release();
        h->pinned_in[i].release();
        h->pinned_out[i].release();
    }
    if (h->small_stream) cudaStreamDestroy(h->small_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->fork_ev) cudaEventDestroy(h->fork_ev);
    if (h->join_ev) cudaEventDestroy(h->join_ev);
    delete h->pool;
    for (auto &b : h->chunk_packed) b.release();
    for (auto &s : h->pipe_stream) if (s) cudaStreamDestroy(s);
    for (auto &e : h->pipe_event) if (e) cudaEventDestroy(e);
    for (auto &e : h->phase_ev) if (e) cudaEventDestroy(e);
    delete h;
    return TM_OK;
}

int tm_prepare_cylinders(tm_handle *h, const float *start, int64_t s_rs, int64_t s_cs, const float *end, int64_t e_rs,
                         int64_t e_cs, int64_t m, float axis_eps, int32_t norm_fma, float *out_len, float *out_unit,
                         void *stream) {
    if (!h) return TM_ERR_INVALID;
    if (m < 0 || (m > 0 && (!start || !end || !out_len || !out_unit)))
        return fail(h, TM_ERR_INVALID, "tm_prepare_cylinders: null pointer or negative size%s%s");
    if (m == 0) return TM_OK;
    TM_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = static_cast<int>((m + 127) / 128);
    if (norm_fma) prepare_kernel<true><<<blocks, 128, 0, st>>>(start, s_rs, s_cs, end, e_rs, e_cs, m, axis_eps, out_len, out_unit);
    else prepare_kernel<false><<<blocks, 128, 0, st>>>(start, s_rs, s_cs, end, e_rs, e_cs, m, axis_eps, out_len, out_unit);
    TM_CUDA(h, cudaGetLastError());
    return TM_OK;
}

int tm_set_cylinders(tm_handle *h, const float *start, int64_t s_rs, int64_t s_cs, const float *unit, int64_t u_rs, int64_t u_cs,
                     const float *length, int64_t l_s, const float *radius, int64_t r_s, const int32_t *ids, int64_t i_s,
                     int64_t m, void *stream) {
    if (!h) return TM_ERR_INVALID;
    if (m < 0 || m > 0x7ffffff0LL) return fail(h, TM_ERR_INVALID, "tm_set_cylinders: cylinder count out of range%s%s");
    if (m > 0 && (!start || !unit || !length || !radius))
        return fail(h, TM_ERR_INVALID, "tm_set_cylinders: null cylinder array%s%s");
    TM_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    h->m = m;
    h->have_cyl = true;
    h->have_grid = false;
    h->n_special = h->n_long = h->n_listed = h->n_aligned = 0;
    if (m == 0) return TM_OK;
    const size_t mm = static_cast<size_t>(m);
    TM_CUDA(h, h->recA.ensure(sizeof(float4) * mm));
    TM_CUDA(h, h->recB.ensure(sizeof(float4) * mm));
    TM_CUDA(h, h->recAB.ensure(sizeof(float4) * 2 * mm));
    TM_CUDA(h, h->ids.ensure(sizeof(int32_t) * mm));
    TM_CUDA(h, h->boxlo.ensure(sizeof(float4) * mm));
    TM_CUDA(h, h->boxhi.ensure(sizeof(float4) * mm));
    TM_CUDA(h, h->special.ensure(sizeof(int32_t) * mm));
    TM_CUDA(h, h->aligned.ensure(sizeof(int32_t) * mm));
    TM_CUDA(h, h->bbox.ensure(sizeof(int) * 14));
    // [0..2] min corner, [3..5] max corner (ordered ints), [6] special, [7] regular, [8] axis-parallel
    const int init[14] = {0x7fffffff, 0x7fffffff, 0x7fffffff, static_cast<int>(0x80000000), static_cast<int>(0x80000000),
                          static_cast<int>(0x80000000), 0, 0, 0, 0, 0, 0, 0, 0};
    TM_CUDA(h, cudaMemcpyAsync(h->bbox.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    const int blocks = static_cast<int>((m + 127) / 128);
    pack_kernel<<<blocks, 128, 0, st>>>(start, s_rs, s_cs, unit, u_rs, u_cs, length, l_s, radius, r_s, ids, i_s,
                                        static_cast<int>(m), h->recA.as<float4>(), h->recB.as<float4>(), h->recAB.as<float4>(),
                                        h->ids.as<int32_t>(),
                                        h->boxlo.as<float4>(), h->boxhi.as<float4>(), h->bbox.as<int>(),
                                        h->special.as<int32_t>(), h->aligned.as<int32_t>());
    TM_CUDA(h, cudaGetLastError());
    int stat[14];
    TM_CUDA(h, cudaMemcpyAsync(stat, h->bbox.p, sizeof(stat), cudaMemcpyDeviceToHost, st));
    TM_CUDA(h, cudaStreamSynchronize(st));      // `init` is a stack buffer; also surfaces bad input pointers here
    unsigned long long len_sum, dia_sum;
    memcpy(&len_sum, stat + 10, sizeof(len_sum));
    memcpy(&dia_sum, stat + 12, sizeof(dia_sum));
    // what sets the number of cylinder pieces per voxel is the spacing along the branches, i.e. the cylinder LENGTH (a thick
    // trunk and a twig of the same length load a voxel alike); the diameter only stands in for tables of zero-length pieces
    const double mean_len = stat[7] > 0 ? static_cast<double>(len_sum) / 1048576.0 / stat[7] : 0.0;
    const double mean_dia = stat[7] > 0 ? static_cast<double>(dia_sum) / 1048576.0 / stat[7] : 0.0;
    h->mean_extent = static_cast<float>(mean_len > 0.0 ? mean_len : mean_dia);
    return TM_OK;
}

static int check_params(tm_handle *h, const tm_params *p) {
    if (!p) return fail(h, TM_ERR_INVALID, "null tm_params%s%s");
    if (p->mode != TM_MODE_AUTO && p->mode != TM_MODE_BRUTE && p->mode != TM_MODE_GRID)
        return fail(h, TM_ERR_INVALID, "tm_params.mode must be TM_MODE_AUTO, TM_MODE_BRUTE or TM_MODE_GRID%s%s");
    if (!(p->perp_atol >= 0.f) || !(p->norm_eps >= 0.f) || !(p->cell_size >= 0.f))
        return fail(h, TM_ERR_INVALID, "tm_params: perp_atol, norm_eps and cell_size must be >= 0%s%s");
    if (p->reserved[0] != 0 || p->reserved[1] != 0) return fail(h, TM_ERR_INVALID, "tm_params.reserved must be 0%s%s");
    return TM_OK;
}

static int label_dispatch(tm_handle *h, const LabelArgs &a) {
    int mode = a.prm.mode;
    if (mode == TM_MODE_AUTO) {
        // the grid costs ~10 launches plus binning; exhaustive search wins while N*M is small
        const double pairs = static_cast<double>(a.n) * static_cast<double>(h->m);
        mode = (pairs <= 3.0e7 || h->m <= 48) ? TM_MODE_BRUTE : TM_MODE_GRID;
    }
    if (mode == TM_MODE_BRUTE) {
        h->stats.mode_used = TM_MODE_BRUTE;
        return label_brute(h, a);
    }
    return label_grid(h, a);
}

int tm_label_points(tm_handle *h, const float *pts, int64_t n, int64_t row_stride, const tm_params *params, int32_t *out_index,
                    int32_t *out_id, float *out_dist, float *out_offset, float *out_radius, void *stream) {
    if (!h) return TM_ERR_INVALID;
    int rc = check_params(h, params);
    if (rc != TM_OK) return rc;
    if (n < 0) return fail(h, TM_ERR_INVALID, "tm_label_points: negative point count%s%s");
    if (n == 0) return TM_OK;
    if (!pts || row_stride < 3) return fail(h, TM_ERR_INVALID, "tm_label_points: null points or row_stride < 3%s%s");
    if (!h->have_cyl) return fail(h, TM_ERR_STATE, "tm_label_points called before tm_set_cylinders%s%s");
    if (h->m == 0) return fail(h, TM_ERR_NO_CYLINDERS, "%s%s", tm_status_string(TM_ERR_NO_CYLINDERS));
    TM_CUDA(h, cudaSetDevice(h->device));
    h->stats = tm_stats{};
    LabelArgs a{pts, n, row_stride, *params, out_index, out_id, out_dist, out_offset, out_radius,
                static_cast<cudaStream_t>(stream)};
    h->last_n = n;
    for (bool &b : h->phase_hit) b = false;
    mark(h, 0, a.stream);
    rc = label_dispatch(h, a);
    mark(h, TM_PHASES - 1, a.stream);
    return rc;
}

int tm_assemble_records(tm_handle *h, const void *cloud, int32_t dtype, int64_t n, int64_t row_stride, const float *offset,
                        const int32_t *id, double *out, void *stream) {
    if (!h) return TM_ERR_INVALID;
    if (n < 0) return fail(h, TM_ERR_INVALID, "tm_assemble_records: negative point count%s%s");
    if (n == 0) return TM_OK;
    if (!cloud || !offset || !id || !out || row_stride < 3 || (dtype != TM_F32 && dtype != TM_F64))
        return fail(h, TM_ERR_INVALID, "tm_assemble_records: bad argument%s%s");
    TM_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = static_cast<int>(std::min<int64_t>((n * 7 + 255) / 256, static_cast<int64_t>(h->sm_count) * 32));
    if (dtype == TM_F32) assemble_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float *>(cloud), n, row_stride, offset, id, out);
    else assemble_kernel<double><<<blocks, 256, 0, st>>>(static_cast<const double *>(cloud), n, row_stride, offset, id, out);
    TM_CUDA(h, cudaGetLastError());
    return TM_OK;
}

// ---- pipelined host entry -----------------------------------------------------------------------
// Staging copies between the caller's pageable arrays and the pinned buffers.  A fresh numpy output array is untouched
// memory: one thread page-faults it at ~7 GB/s, which is slower than the PCIe link, so large copies are split over a
// few host threads (bounded by the CPUs this process may run on).
static unsigned host_copy_threads() {
    static const unsigned n = [] {
        cpu_set_t set;
        unsigned avail = 0;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) avail = static_cast<unsigned>(CPU_COUNT(&set));
        if (avail == 0) avail = std::thread::hardware_concurrency();
        return std::max(1u, std::min(avail, 8u));
    }();
    return n;
}

static void par_memcpy(void *dst, const void *src, size_t bytes) {
    const unsigned nt = host_copy_threads();
    if (bytes < (8u << 20) || nt <= 1) { memcpy(dst, src, bytes); return; }
    const size_t per = ((bytes / nt) + 4095) & ~static_cast<size_t>(4095);
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (unsigned t = 1; t < nt; ++t) {
        const size_t lo = t * per;
        if (lo >= bytes) break;
        const size_t len = std::min(per, bytes - lo);
        pool.emplace_back([=] { memcpy(static_cast<unsigned char *>(dst) + lo, static_cast<const unsigned char *>(src) + lo, len); });
    }
    memcpy(dst, src, std::min(per, bytes));
    for (auto &th : pool) th.join();
}

static bool host_is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

static unsigned host_threads_available() {
    cpu_set_t set;
    unsigned avail = 0;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) avail = static_cast<unsigned>(CPU_COUNT(&set));
    if (avail == 0) avail = std::thread::hardware_concurrency();
    // one process per GPU (torchrun): the ranks of a node share the host cores
    if (const char *lws = getenv("LOCAL_WORLD_SIZE")) { const int r = atoi(lws); if (r > 1) avail /= static_cast<unsigned>(r); }
    return std::max(1u, avail);
}

}  // extern "C" (reopened below)

// (N,7) float64 rows [x, y, z, ox, oy, oz, ID] (LabelGenerationCuda.py:114,131-133) written by the HOST: xyz come from
// the caller's own cloud at its precision, only the 16-byte {offset, id} records cross PCIe.  Streaming stores: the
// record array is written once and never read here.
static inline void store_f64(double *dst, double v) {
#if defined(__x86_64__)
    long long bits;
    memcpy(&bits, &v, 8);
    _mm_stream_si64(reinterpret_cast<long long *>(dst), bits);
#else
    *dst = v;
#endif
}

template <typename T>
static void assemble_rows_host(const T *cloud, int64_t row_stride, const float4 *packed, double *out, int64_t r0, int64_t r1,
                               int32_t width, const double *tail) {
    for (int64_t r = r0; r < r1; ++r) {
        const T *p = cloud + r * row_stride;
        const float4 k = packed[r];
        int32_t id;
        memcpy(&id, &k.w, 4);
        double *o = out + static_cast<int64_t>(width) * r;
        store_f64(o + 0, static_cast<double>(p[0]));
        store_f64(o + 1, static_cast<double>(p[1]));
        store_f64(o + 2, static_cast<double>(p[2]));
        store_f64(o + 3, static_cast<double>(k.x));
        store_f64(o + 4, static_cast<double>(k.y));
        store_f64(o + 5, static_cast<double>(k.z));
        store_f64(o + 6, static_cast<double>(id));
        for (int32_t c = 7; c < width; ++c) store_f64(o + c, tail[c - 7]);       // the drivers' feature columns (ones)
    }
#if defined(__x86_64__)
    _mm_sfence();
#endif
}

// Staging of a chunk for the H2D copy: the rows' xyz as compact float32 (n,3) in page-locked memory.  float64 clouds are
// rounded here (RNE, what torch.tensor(points, dtype=torch.float32) does at LabelGenerationCuda.py:33), so 12 bytes per
// point cross PCIe whatever the caller's dtype or row stride.
template <typename T>
static void stage_rows_f32(const T *cloud, int64_t row_stride, float *dst, int64_t r0, int64_t r1) {
    for (int64_t r = r0; r < r1; ++r) {
        const T *p = cloud + r * row_stride;
        dst[3 * r + 0] = static_cast<float>(p[0]);
        dst[3 * r + 1] = static_cast<float>(p[1]);
        dst[3 * r + 2] = static_cast<float>(p[2]);
    }
}

// tm_label_cloud_host, host-assembly variant: per chunk H2D -> label (packed {offset, id}) -> D2H 16 B/point, while the
// host workers assemble earlier chunks' records.  Up to PIPE_SLOTS chunks are in flight; a chunk's buffers are reused
// only after the chunk has been assembled (hence fully transferred).
static int label_cloud_host_assemble(tm_handle *h, const void *cloud_host, int32_t dtype, int64_t n, int64_t row_stride,
                                     const tm_params *params, double *out_records_host, float *out_dist_host, unsigned nthreads,
                                     int32_t width = 7, const double *tail = nullptr) {
    if (!h->pool || h->pool->n != nthreads) {
        delete h->pool;
        h->pool = new (std::nothrow) tmn::HostPool(nthreads);
        if (!h->pool) return tmn::fail(h, TM_ERR_NOMEM, "host worker pool%s%s");
    }
    cudaStream_t s_in = h->pipe_stream[0], s_cmp = h->pipe_stream[1], s_out = h->pipe_stream[2], s_rows = h->pipe_stream[3];
    h->host_assembly_threads = static_cast<int32_t>(nthreads);
    // About four chunks per call: every chunk costs the calling thread ~0.15 ms of fixed work (a dozen launches, two copies,
    // two hand-offs to the worker pool), and with fewer than three the copies no longer hide behind the host's passes
    // (10M points: 20 chunks 15.2 ms, 10 chunks 13.2, 5 chunks 12.4, 3 chunks 12.2, 2 chunks 14.0 —
    // profiles/r02_e2e_chunks.json).  Bounded below (small clouds: the fixed work) and above (page-locked staging: 32 bytes
    // per point and slot).
    int64_t chunk = std::min<int64_t>(std::max<int64_t>((n + 3) / 4, 512 << 10), 2560 << 10);
    chunk = (chunk + 4095) & ~static_cast<int64_t>(4095);
    if (const char *env = getenv("TM_HOST_CHUNK")) { const long long v = atoll(env); if (v >= 1024) chunk = v; }
    chunk = std::min(chunk, n);
    int depth = tmn::PIPE_SLOTS;
    if (const char *env = getenv("TM_HOST_DEPTH")) depth = std::max(2, std::min(tmn::PIPE_SLOTS, atoi(env)));
    const size_t esz = dtype == TM_F32 ? 4 : 8;
    // a page-locked compact float32 cloud is copied from where it lies; everything else (pageable memory, float64, padded
    // rows) is staged as compact float32 by the host workers
    const bool direct_in = host_is_pinned(cloud_host) && dtype == TM_F32 && row_stride == 3;
    const size_t in_bytes = static_cast<size_t>(chunk) * 3 * sizeof(float);
    const size_t pk_bytes = static_cast<size_t>(chunk) * sizeof(float4);
    const size_t out_bytes = pk_bytes + static_cast<size_t>(chunk) * sizeof(float);
    // A pinned record array can also be written by the copy engine: TM_HOST_SPLIT percent of the chunks are assembled on
    // the device and DMA'd straight into the caller's rows (56 B/point over PCIe, on their own stream) while the host
    // workers assemble the others (16 B/point + the host's stores).  See profiles/r01i_host_pipeline.md.  (float32 clouds
    // only: the device sees the float32 rounding of a float64 cloud, the records carry the caller's own values.)
    int split = 0;
    if (const char *env = getenv("TM_HOST_SPLIT")) { if (width == 7 && dtype == TM_F32 && host_is_pinned(out_records_host)) split = std::max(0, std::min(100, atoi(env))); }
    auto on_device = [split](int64_t c) { return ((c + 1) * split) / 100 > (c * split) / 100; };
    for (int b = 0; b < depth; ++b) {
        if (!direct_in) TM_CUDA(h, h->pinned_in[b].ensure(in_bytes));
        TM_CUDA(h, h->pinned_out[b].ensure(out_bytes));
        TM_CUDA(h, h->chunk_in[b].ensure(in_bytes));
        TM_CUDA(h, h->chunk_packed[b].ensure(pk_bytes));
        TM_CUDA(h, h->chunk_dist[b].ensure(static_cast<size_t>(chunk) * 4));
        if (split) {
            TM_CUDA(h, h->chunk_rec[b].ensure(static_cast<size_t>(chunk) * 7 * sizeof(double)));
            TM_CUDA(h, h->chunk_off[b].ensure(static_cast<size_t>(chunk) * 12));
            TM_CUDA(h, h->chunk_id[b].ensure(static_cast<size_t>(chunk) * 4));
        }
    }

    // events per slot b: [b] H2D done, [SLOTS + b] label done, [2 SLOTS + b] D2H done
    constexpr int EV_H2D = 0, EV_CMP = tmn::PIPE_SLOTS, EV_D2H = 2 * tmn::PIPE_SLOTS;
    tm_stats total{};
    size_t d2h_total = 0;
    const bool trace = getenv("TM_TRACE_HOST") != nullptr;
    double t_issue = 0, t_wait = 0, t_asm = 0, t_stage = 0;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    std::vector<cudaEvent_t> tev;                       // trace only: [4c] H2D start, [4c+1] H2D end, [4c+2] label end, [4c+3] D2H end
    auto tmark = [&](cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
    const int64_t nchunks = (n + chunk - 1) / chunk;
    const unsigned char *src = static_cast<const unsigned char *>(cloud_host);

    auto issue = [&](int64_t c) -> int {
        const int b = static_cast<int>(c % depth);
        const int64_t cnt = std::min(chunk, n - c * chunk);
        const size_t bytes = static_cast<size_t>(cnt) * 3 * sizeof(float);
        const unsigned char *csrc = src + static_cast<size_t>(c) * static_cast<size_t>(chunk) * static_cast<size_t>(row_stride) * esz;
        const double i0 = now();
        if (!direct_in) {
            float *dst = static_cast<float *>(h->pinned_in[b].p);
            h->pool->run([=](unsigned t, unsigned nt) {
                const int64_t per = (cnt + nt - 1) / nt, r0 = std::min<int64_t>(cnt, t * per), r1 = std::min<int64_t>(cnt, r0 + per);
                if (dtype == TM_F32) stage_rows_f32(reinterpret_cast<const float *>(csrc), row_stride, dst, r0, r1);
                else stage_rows_f32(reinterpret_cast<const double *>(csrc), row_stride, dst, r0, r1);
            });
            t_stage += now() - i0;
        }
        tmark(s_in);
        TM_CUDA(h, cudaMemcpyAsync(h->chunk_in[b].p, direct_in ? static_cast<const void *>(csrc) : h->pinned_in[b].p, bytes,
                                   cudaMemcpyHostToDevice, s_in));
        tmark(s_in);
        TM_CUDA(h, cudaEventRecord(h->pipe_event[EV_H2D + b], s_in));
        TM_CUDA(h, cudaStreamWaitEvent(s_cmp, h->pipe_event[EV_H2D + b], 0));
        const float *pts32 = h->chunk_in[b].as<float>();
        const int64_t stride32 = 3;
        h->stats = tm_stats{};
        const bool dev = on_device(c);
        LabelArgs a{pts32, cnt, stride32, *params, nullptr, dev ? h->chunk_id[b].as<int32_t>() : nullptr,
                    out_dist_host ? h->chunk_dist[b].as<float>() : nullptr, dev ? h->chunk_off[b].as<float>() : nullptr, nullptr, s_cmp};
        if (!dev) a.out_packed = h->chunk_packed[b].as<float4>();
        int rc = label_dispatch(h, a);
        if (rc != TM_OK) return rc;
        total.pairs_evaluated += h->stats.pairs_evaluated;
        total.points_brute += h->stats.points_brute;
        h->last_n = cnt;
        if (dev) {
            rc = tm_assemble_records(h, h->chunk_in[b].p, TM_F32, cnt, 3, h->chunk_off[b].as<float>(),
                                     h->chunk_id[b].as<int32_t>(), h->chunk_rec[b].as<double>(), s_cmp);
            if (rc != TM_OK) return rc;
        }
        tmark(s_cmp);
        TM_CUDA(h, cudaEventRecord(h->pipe_event[EV_CMP + b], s_cmp));
        cudaStream_t so = dev ? s_rows : s_out;
        TM_CUDA(h, cudaStreamWaitEvent(so, h->pipe_event[EV_CMP + b], 0));
        if (dev) {
            TM_CUDA(h, cudaMemcpyAsync(out_records_host + c * chunk * 7, h->chunk_rec[b].p, static_cast<size_t>(cnt) * 7 * sizeof(double),
                                       cudaMemcpyDeviceToHost, so));
            d2h_total += static_cast<size_t>(cnt) * 56;
        } else {
            TM_CUDA(h, cudaMemcpyAsync(h->pinned_out[b].p, h->chunk_packed[b].p, static_cast<size_t>(cnt) * sizeof(float4),
                                       cudaMemcpyDeviceToHost, so));
            d2h_total += static_cast<size_t>(cnt) * 16;
        }
        if (out_dist_host)
            TM_CUDA(h, cudaMemcpyAsync(static_cast<unsigned char *>(h->pinned_out[b].p) + pk_bytes, h->chunk_dist[b].p,
                                       static_cast<size_t>(cnt) * sizeof(float), cudaMemcpyDeviceToHost, so));
        tmark(so);
        TM_CUDA(h, cudaEventRecord(h->pipe_event[EV_D2H + b], so));
        t_issue += now() - i0;
        return TM_OK;
    };
    auto assemble = [&](int64_t c) -> int {
        const int b = static_cast<int>(c % depth);
        const int64_t cnt = std::min(chunk, n - c * chunk);
        const double w0 = now();
        TM_CUDA(h, cudaEventSynchronize(h->pipe_event[EV_D2H + b]));
        const double w1 = now();
        t_wait += w1 - w0;
        if (!on_device(c)) {
            const float4 *packed = static_cast<const float4 *>(h->pinned_out[b].p);
            const unsigned char *crow = src + static_cast<size_t>(c) * static_cast<size_t>(chunk) * static_cast<size_t>(row_stride) * esz;
            double *orow = out_records_host + c * chunk * width;
            double tl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int32_t k = 7; k < width; ++k) tl[k - 7] = tail[k - 7];
            h->pool->run([=](unsigned t, unsigned nt) {
                const int64_t per = (cnt + nt - 1) / nt, r0 = std::min<int64_t>(cnt, t * per), r1 = std::min<int64_t>(cnt, r0 + per);
                if (dtype == TM_F32) assemble_rows_host(reinterpret_cast<const float *>(crow), row_stride, packed, orow, r0, r1, width, tl);
                else assemble_rows_host(reinterpret_cast<const double *>(crow), row_stride, packed, orow, r0, r1, width, tl);
            });
        }
        if (out_dist_host)
            memcpy(out_dist_host + c * chunk, static_cast<unsigned char *>(h->pinned_out[b].p) + pk_bytes, static_cast<size_t>(cnt) * sizeof(float));
        t_asm += now() - w1;
        return TM_OK;
    };

    int64_t issued = 0;
    for (int64_t done = 0; done < nchunks; ++done) {
        while (issued < nchunks && issued - done < depth) {
            const int rc = issue(issued++);
            if (rc != TM_OK) { cudaDeviceSynchronize(); return rc; }
        }
        const int rc = assemble(done);
        if (rc != TM_OK) { cudaDeviceSynchronize(); return rc; }
    }
    TM_CUDA(h, cudaStreamSynchronize(s_cmp));
    h->stats.pairs_evaluated = total.pairs_evaluated;
    h->stats.points_brute = total.points_brute;
    if (trace) {
        float h2d = 0, lab = 0, d2h = 0, ms = 0;
        for (size_t c = 0; c + 3 < tev.size(); c += 4) {
            cudaEventElapsedTime(&ms, tev[c], tev[c + 1]); h2d += ms;
            cudaEventElapsedTime(&ms, tev[c + 1], tev[c + 2]); lab += ms;
            cudaEventElapsedTime(&ms, tev[c + 2], tev[c + 3]); d2h += ms;
        }
        fprintf(stderr, "[tm host] device side, summed over chunks: H2D %.3f ms, H2D end -> label end %.3f, label end -> D2H end %.3f\n", h2d, lab, d2h);
        for (auto e : tev) cudaEventDestroy(e);
        fprintf(stderr, "[tm host] %lld chunks of %lld, %d in flight: issue %.3f ms (of which staging %.3f), wait for D2H %.3f, assemble %.3f\n",
                static_cast<long long>(nchunks), static_cast<long long>(chunk), depth, t_issue, t_stage, t_wait, t_asm);
    }
    h->host_d2h_bytes_per_point = static_cast<int32_t>((d2h_total + static_cast<size_t>(n) / 2) / static_cast<size_t>(n)) + (out_dist_host ? 4 : 0);
    return TM_OK;
}

extern "C" {


int tm_label_cloud_host(tm_handle *h, const void *cloud_host, int32_t dtype, int64_t n, int64_t row_stride,
                        const tm_params *params, double *out_records_host, float *out_dist_host) {
    if (!h) return TM_ERR_INVALID;
    int rc = check_params(h, params);
    if (rc != TM_OK) return rc;
    if (n < 0) return fail(h, TM_ERR_INVALID, "tm_label_cloud_host: negative point count%s%s");
    if (n == 0) return TM_OK;
    if (!cloud_host || !out_records_host || row_stride < 3 || (dtype != TM_F32 && dtype != TM_F64))
        return fail(h, TM_ERR_INVALID, "tm_label_cloud_host: bad argument%s%s");
    if (!h->have_cyl) return fail(h, TM_ERR_STATE, "tm_label_cloud_host called before tm_set_cylinders%s%s");
    if (h->m == 0) return fail(h, TM_ERR_NO_CYLINDERS, "%s%s", tm_status_string(TM_ERR_NO_CYLINDERS));
    TM_CUDA(h, cudaSetDevice(h->device));
    for (auto &s : h->pipe_stream) if (!s) TM_CUDA(h, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (auto &e : h->pipe_event) if (!e) TM_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaStream_t s_in = h->pipe_stream[0], s_cmp = h->pipe_stream[1], s_out = h->pipe_stream[2];

    // With a few host threads to spare the records are assembled on the host: only 16 bytes per point come back over PCIe
    // instead of 56, and the host writes the xyz it already has.  TM_HOST_ASSEMBLE=0 forces the device-assembled path.
    {
        const char *env = getenv("TM_HOST_ASSEMBLE");
        const int want = env ? atoi(env) : -1;
        const unsigned ht = std::min(host_threads_available(), 16u);
        if (want != 0 && (ht >= 4 || want > 0))
            return label_cloud_host_assemble(h, cloud_host, dtype, n, row_stride, params, out_records_host, out_dist_host,
                                             want > 0 ? std::max(1u, std::min(static_cast<unsigned>(want), 64u)) : ht);
    }

    h->host_d2h_bytes_per_point = 56 + (out_dist_host ? 4 : 0);
    h->host_assembly_threads = 0;
    int64_t chunk = 1 << 20;
    if (const char *env = getenv("TM_HOST_CHUNK")) { const long long v = atoll(env); if (v >= 1024) chunk = v; }
    chunk = std::min(chunk, n);
    const size_t esz = dtype == TM_F32 ? 4 : 8;
    const bool in_pinned = host_is_pinned(cloud_host);
    const bool out_pinned = host_is_pinned(out_records_host) && (!out_dist_host || host_is_pinned(out_dist_host));
    const size_t in_bytes = static_cast<size_t>(chunk) * static_cast<size_t>(row_stride) * esz;
    const size_t rec_bytes = static_cast<size_t>(chunk) * 7 * sizeof(double);
    const size_t out_bytes = rec_bytes + static_cast<size_t>(chunk) * sizeof(float);
    for (int b = 0; b < 2; ++b) {
        if (!in_pinned) TM_CUDA(h, h->pinned_in[b].ensure(in_bytes));
        if (!out_pinned) TM_CUDA(h, h->pinned_out[b].ensure(out_bytes));
        TM_CUDA(h, h->chunk_in[b].ensure(in_bytes + (dtype == TM_F64 ? static_cast<size_t>(chunk) * 12 : 0)));
        TM_CUDA(h, h->chunk_rec[b].ensure(rec_bytes));
        TM_CUDA(h, h->chunk_off[b].ensure(static_cast<size_t>(chunk) * 12));
        TM_CUDA(h, h->chunk_id[b].ensure(static_cast<size_t>(chunk) * 4));
        TM_CUDA(h, h->chunk_dist[b].ensure(static_cast<size_t>(chunk) * 4));
    }

    // events: [0,1] H2D done per buffer, [2,3] compute done, [4,5] D2H done, [6,7] compute consumed input
    tm_stats total{};
    const int64_t nchunks = (n + chunk - 1) / chunk;
    const unsigned char *src = static_cast<const unsigned char *>(cloud_host);
    auto drain_out = [&](int64_t c) -> int {           // copy a finished chunk from pinned staging to the caller
        const int b = static_cast<int>(c & 1);
        TM_CUDA(h, cudaEventSynchronize(h->pipe_event[4 + b]));
        if (!out_pinned) {
            const int64_t cnt = std::min(chunk, n - c * chunk);
            par_memcpy(out_records_host + c * chunk * 7, h->pinned_out[b].p, static_cast<size_t>(cnt) * 7 * sizeof(double));
            if (out_dist_host)
                memcpy(out_dist_host + c * chunk, static_cast<unsigned char *>(h->pinned_out[b].p) + rec_bytes,
                       static_cast<size_t>(cnt) * sizeof(float));
        }
        return TM_OK;
    };
    // an error leaves copies of earlier chunks in flight into the caller's (or the staging) memory: wait for them before
    // handing control back
    auto bail = [&](int code) { cudaDeviceSynchronize(); return code; };
#define TM_CUDA_BAIL(expr)                                                                                          \
    do {                                                                                                            \
        cudaError_t _e = (expr);                                                                                    \
        if (_e != cudaSuccess) {                                                                                    \
            cudaDeviceSynchronize();                                                                                \
            return tmn::fail(h, _e == cudaErrorMemoryAllocation ? TM_ERR_NOMEM : TM_ERR_CUDA, "%s failed: %s", #expr, \
                             cudaGetErrorString(_e));                                                               \
        }                                                                                                           \
    } while (0)
    for (int64_t c = 0; c < nchunks; ++c) {
        const int b = static_cast<int>(c & 1);
        const int64_t cnt = std::min(chunk, n - c * chunk);
        const size_t bytes = static_cast<size_t>(cnt) * static_cast<size_t>(row_stride) * esz;
        const unsigned char *csrc = src + static_cast<size_t>(c) * static_cast<size_t>(chunk) * static_cast<size_t>(row_stride) * esz;
        if (c >= 2) {
            // buffer b is being reused: its previous output must have left the device and the staging area
            rc = drain_out(c - 2);
            if (rc != TM_OK) return bail(rc);
            TM_CUDA_BAIL(cudaStreamWaitEvent(s_in, h->pipe_event[6 + b], 0));
        }
        if (!in_pinned) {
            par_memcpy(h->pinned_in[b].p, csrc, bytes);
            TM_CUDA_BAIL(cudaMemcpyAsync(h->chunk_in[b].p, h->pinned_in[b].p, bytes, cudaMemcpyHostToDevice, s_in));
        } else {
            TM_CUDA_BAIL(cudaMemcpyAsync(h->chunk_in[b].p, csrc, bytes, cudaMemcpyHostToDevice, s_in));
        }
        TM_CUDA_BAIL(cudaEventRecord(h->pipe_event[0 + b], s_in));
        TM_CUDA_BAIL(cudaStreamWaitEvent(s_cmp, h->pipe_event[0 + b], 0));
        if (c >= 2) TM_CUDA_BAIL(cudaStreamWaitEvent(s_cmp, h->pipe_event[4 + b], 0));   // records buffer b free again
        const float *pts32;
        int64_t stride32;
        if (dtype == TM_F64) {
            float *conv = reinterpret_cast<float *>(h->chunk_in[b].as<unsigned char>() + in_bytes);
            const int blocks = static_cast<int>(std::min<int64_t>((cnt * 3 + 255) / 256, static_cast<int64_t>(h->sm_count) * 32));
            f64_to_f32_xyz_kernel<<<blocks, 256, 0, s_cmp>>>(h->chunk_in[b].as<double>(), cnt, row_stride, conv);
            TM_CUDA_BAIL(cudaGetLastError());
            pts32 = conv;
            stride32 = 3;
        } else {
            pts32 = h->chunk_in[b].as<float>();
            stride32 = row_stride;
        }
        h->stats = tm_stats{};
        LabelArgs a{pts32, cnt, stride32, *params, nullptr, h->chunk_id[b].as<int32_t>(),
                    out_dist_host ? h->chunk_dist[b].as<float>() : nullptr, h->chunk_off[b].as<float>(), nullptr, s_cmp};
        rc = label_dispatch(h, a);
        if (rc != TM_OK) return bail(rc);
        total.pairs_evaluated += h->stats.pairs_evaluated;
        total.points_brute += h->stats.points_brute;
        total.mode_used = h->stats.mode_used;
        h->last_n = cnt;
        rc = tm_assemble_records(h, h->chunk_in[b].p, dtype, cnt, row_stride, h->chunk_off[b].as<float>(),
                                 h->chunk_id[b].as<int32_t>(), h->chunk_rec[b].as<double>(), s_cmp);
        if (rc != TM_OK) return bail(rc);
        TM_CUDA_BAIL(cudaEventRecord(h->pipe_event[2 + b], s_cmp));
        TM_CUDA_BAIL(cudaEventRecord(h->pipe_event[6 + b], s_cmp));
        TM_CUDA_BAIL(cudaStreamWaitEvent(s_out, h->pipe_event[2 + b], 0));
        double *dst_rec = out_pinned ? out_records_host + c * chunk * 7 : static_cast<double *>(h->pinned_out[b].p);
        TM_CUDA_BAIL(cudaMemcpyAsync(dst_rec, h->chunk_rec[b].p, static_cast<size_t>(cnt) * 7 * sizeof(double),
                                   cudaMemcpyDeviceToHost, s_out));
        if (out_dist_host) {
            float *dst_d = out_pinned ? out_dist_host + c * chunk
                                      : reinterpret_cast<float *>(static_cast<unsigned char *>(h->pinned_out[b].p) + rec_bytes);
            TM_CUDA_BAIL(cudaMemcpyAsync(dst_d, h->chunk_dist[b].p, static_cast<size_t>(cnt) * sizeof(float),
                                       cudaMemcpyDeviceToHost, s_out));
        }
        TM_CUDA_BAIL(cudaEventRecord(h->pipe_event[4 + b], s_out));
    }
    for (int64_t c = std::max<int64_t>(0, nchunks - 2); c < nchunks; ++c) {
        rc = drain_out(c);
        if (rc != TM_OK) return bail(rc);
    }
#undef TM_CUDA_BAIL
    TM_CUDA(h, cudaStreamSynchronize(s_out));
    TM_CUDA(h, cudaStreamSynchronize(s_cmp));
    h->stats.pairs_evaluated = total.pairs_evaluated;
    h->stats.points_brute = total.points_brute;
    return TM_OK;
}

int tm_label_cloud_host_wide(tm_handle *h, const void *cloud_host, int32_t dtype, int64_t n, int64_t row_stride,
                             const tm_params *params, double *out_rows_host, int32_t row_doubles, const double *tail_values,
                             float *out_dist_host) {
    if (!h) return TM_ERR_INVALID;
    if (row_doubles == 7) return tm_label_cloud_host(h, cloud_host, dtype, n, row_stride, params, out_rows_host, out_dist_host);
    int rc = check_params(h, params);
    if (rc != TM_OK) return rc;
    if (n < 0) return fail(h, TM_ERR_INVALID, "tm_label_cloud_host_wide: negative point count%s%s");
    if (n == 0) return TM_OK;
    if (!cloud_host || !out_rows_host || row_stride < 3 || (dtype != TM_F32 && dtype != TM_F64) || row_doubles < 7 ||
        row_doubles > 15 || !tail_values)
        return fail(h, TM_ERR_INVALID, "tm_label_cloud_host_wide: bad argument%s%s");
    if (!h->have_cyl) return fail(h, TM_ERR_STATE, "tm_label_cloud_host_wide called before tm_set_cylinders%s%s");
    if (h->m == 0) return fail(h, TM_ERR_NO_CYLINDERS, "%s%s", tm_status_string(TM_ERR_NO_CYLINDERS));
    TM_CUDA(h, cudaSetDevice(h->device));
    for (auto &s : h->pipe_stream) if (!s) TM_CUDA(h, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (auto &e : h->pipe_event) if (!e) TM_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // the wide rows are always written by the host workers (the device never sees the feature columns)
    const unsigned ht = std::max(1u, std::min(host_threads_available(), 16u));
    return label_cloud_host_assemble(h, cloud_host, dtype, n, row_stride, params, out_rows_host, out_dist_host, ht, row_doubles,
                                     tail_values);
}

// ---- small-table fast path ------------------------------------------------------------------------
int tm_cloud_upload_host(tm_handle *h, const void *cloud_host, int32_t dtype, int64_t n, int64_t row_stride) {
    if (!h) return TM_ERR_INVALID;
    if (n < 0 || (n > 0 && (!cloud_host || row_stride < 3)) || (dtype != TM_F32 && dtype != TM_F64))
        return fail(h, TM_ERR_INVALID, "tm_cloud_upload_host: bad argument%s%s");
    TM_CUDA(h, cudaSetDevice(h->device));
    if (!h->small_stream) TM_CUDA(h, cudaStreamCreateWithFlags(&h->small_stream, cudaStreamNonBlocking));
    cudaStream_t st = h->small_stream;
    h->cloud_res_n = -1;
    if (n == 0) { h->cloud_res_n = 0; return TM_OK; }
    const size_t esz = dtype == TM_F32 ? 4 : 8;
    TM_CUDA(h, h->cloud_res.ensure(sizeof(float) * 3 * static_cast<size_t>(n)));
    if (dtype == TM_F32 && row_stride == 3) {
        TM_CUDA(h, cudaMemcpyAsync(h->cloud_res.p, cloud_host, sizeof(float) * 3 * static_cast<size_t>(n), cudaMemcpyHostToDevice, st));
    } else {
        // wider rows / float64: stage the raw rows, then gather + round to fp32 on the device
        // (torch.tensor(points, dtype=torch.float32): round to nearest even)
        const size_t raw = static_cast<size_t>(n) * static_cast<size_t>(row_stride) * esz;
        TM_CUDA(h, h->chunk_in[0].ensure(raw));
        TM_CUDA(h, cudaMemcpyAsync(h->chunk_in[0].p, cloud_host, raw, cudaMemcpyHostToDevice, st));
        const int blocks = static_cast<int>(std::min<int64_t>((n * 3 + 255) / 256, static_cast<int64_t>(h->sm_count) * 32));
        if (dtype == TM_F64) f64_to_f32_xyz_kernel<<<blocks, 256, 0, st>>>(h->chunk_in[0].as<double>(), n, row_stride, h->cloud_res.as<float>());
        else f32_xyz_kernel<<<blocks, 256, 0, st>>>(h->chunk_in[0].as<float>(), n, row_stride, h->cloud_res.as<float>());
        TM_CUDA(h, cudaGetLastError());
    }
    TM_CUDA(h, cudaStreamSynchronize(st));
    h->cloud_res_n = n;
    return TM_OK;
}

int tm_proximity_flags_host(tm_handle *h, const int64_t *subset_host, int64_t n, const float *start_host, const float *end_host,
                            const float *radius_host, int64_t m, const tm_params *params, float axis_eps, float eps,
                            uint8_t *out_flags_host, float *out_dist_host, int32_t *out_index_host) {
    if (!h) return TM_ERR_INVALID;
    int rc = check_params(h, params);
    if (rc != TM_OK) return rc;
    if (h->cloud_res_n < 0) return fail(h, TM_ERR_STATE, "tm_proximity_flags_host called before tm_cloud_upload_host%s%s");
    if (n < 0 || m < 0) return fail(h, TM_ERR_INVALID, "tm_proximity_flags_host: negative size%s%s");
    if (!subset_host && n > h->cloud_res_n) return fail(h, TM_ERR_INVALID, "tm_proximity_flags_host: more points than the resident cloud holds%s%s");
    if (n == 0) return TM_OK;
    if (m == 0) return fail(h, TM_ERR_NO_CYLINDERS, "%s%s", tm_status_string(TM_ERR_NO_CYLINDERS));
    if (m > tmn::SMALL_MAX_M) return fail(h, TM_ERR_INVALID, "tm_proximity_flags_host: too many cylinders for the small-table path%s%s");
    if (!start_host || !end_host || !radius_host) return fail(h, TM_ERR_INVALID, "tm_proximity_flags_host: null cylinder array%s%s");
    TM_CUDA(h, cudaSetDevice(h->device));
    if (!h->small_stream) TM_CUDA(h, cudaStreamCreateWithFlags(&h->small_stream, cudaStreamNonBlocking));
    cudaStream_t st = h->small_stream;
    if (subset_host)
        for (int64_t i = 0; i < n; ++i)
            if (subset_host[i] < 0 || subset_host[i] >= h->cloud_res_n)
                return fail(h, TM_ERR_INVALID, "tm_proximity_flags_host: subset index outside the resident cloud%s%s");

    // one staging block per direction: [subset rows | cylinders] in, [dist | index | flags] out
    const size_t idx_bytes = subset_host ? sizeof(int64_t) * static_cast<size_t>(n) : 0;
    const size_t cyl_bytes = sizeof(float) * 7 * static_cast<size_t>(m);
    TM_CUDA(h, h->small_in.ensure(idx_bytes + cyl_bytes));
    TM_CUDA(h, h->small_out.ensure(static_cast<size_t>(n) * 9 + 16));
    TM_CUDA(h, h->pinned_in[0].ensure(std::max<size_t>(idx_bytes + cyl_bytes, 1 << 20)));
    unsigned char *stage = static_cast<unsigned char *>(h->pinned_in[0].p);
    if (subset_host) memcpy(stage, subset_host, idx_bytes);
    float *cyl = reinterpret_cast<float *>(stage + idx_bytes);
    for (int64_t c = 0; c < m; ++c) {
        cyl[7 * c + 0] = start_host[3 * c]; cyl[7 * c + 1] = start_host[3 * c + 1]; cyl[7 * c + 2] = start_host[3 * c + 2];
        cyl[7 * c + 3] = end_host[3 * c];   cyl[7 * c + 4] = end_host[3 * c + 1];   cyl[7 * c + 5] = end_host[3 * c + 2];
        cyl[7 * c + 6] = radius_host[c];
    }
    // This caller is latency-bound (thousands of calls per tree on a few thousand rows).  Up to 64k rows the kernel reads
    // the staged rows / cylinders from the page-locked block itself and writes its results into page-locked memory (the
    // host pointers are device-accessible): ONE launch and one synchronisation per call instead of copy + launch + up to
    // three copies into pageable arrays.  Larger subsets go through device buffers with DMA copies.
    const bool direct = n <= 65536 && getenv("TM_SMALL_STAGED") == nullptr;
    tmn::SmallArgs a;
    a.cloud = h->cloud_res.as<float>();
    a.n = n;
    a.m = static_cast<int>(m);
    a.axis_eps = axis_eps; a.atol = params->perp_atol; a.eps_norm = params->norm_eps; a.eps_flag = eps;
    unsigned char *ob;
    if (direct) {
        TM_CUDA(h, h->pinned_out[0].ensure(std::max<size_t>(static_cast<size_t>(n) * 9 + 16, 1 << 20)));
        a.subset = subset_host ? reinterpret_cast<const int64_t *>(stage) : nullptr;
        a.cyl = cyl;
        ob = static_cast<unsigned char *>(h->pinned_out[0].p);
    } else {
        TM_CUDA(h, cudaMemcpyAsync(h->small_in.p, stage, idx_bytes + cyl_bytes, cudaMemcpyHostToDevice, st));
        a.subset = subset_host ? h->small_in.as<int64_t>() : nullptr;
        a.cyl = reinterpret_cast<const float *>(h->small_in.as<unsigned char>() + idx_bytes);
        ob = h->small_out.as<unsigned char>();
    }
    a.dist = out_dist_host ? reinterpret_cast<float *>(ob) : nullptr;
    a.index = out_index_host ? reinterpret_cast<int32_t *>(ob + 4 * static_cast<size_t>(n)) : nullptr;
    a.flags = out_flags_host ? ob + 8 * static_cast<size_t>(n) : nullptr;
    rc = tmn::run_proximity(h, a, params->norm_eps > 0.f, params->norm_fma != 0, st);
    if (rc != TM_OK) return rc;
    if (direct) {
        TM_CUDA(h, cudaStreamSynchronize(st));
        if (out_dist_host) memcpy(out_dist_host, a.dist, 4 * static_cast<size_t>(n));
        if (out_index_host) memcpy(out_index_host, a.index, 4 * static_cast<size_t>(n));
        if (out_flags_host) memcpy(out_flags_host, a.flags, static_cast<size_t>(n));
    } else {
        if (out_dist_host) TM_CUDA(h, cudaMemcpyAsync(out_dist_host, a.dist, 4 * static_cast<size_t>(n), cudaMemcpyDeviceToHost, st));
        if (out_index_host) TM_CUDA(h, cudaMemcpyAsync(out_index_host, a.index, 4 * static_cast<size_t>(n), cudaMemcpyDeviceToHost, st));
        if (out_flags_host) TM_CUDA(h, cudaMemcpyAsync(out_flags_host, a.flags, static_cast<size_t>(n), cudaMemcpyDeviceToHost, st));
        TM_CUDA(h, cudaStreamSynchronize(st));
    }
    h->stats = tm_stats{};
    h->stats.mode_used = TM_MODE_BRUTE;
    h->stats.pairs_evaluated = static_cast<uint64_t>(n) * static_cast<uint64_t>(m);
    h->stats.points_brute = static_cast<uint64_t>(n);
    return TM_OK;
}

int tm_measure_host_bandwidth(tm_handle *h, double *bytes_per_second, int32_t *threads) {
    if (!h || !bytes_per_second) return TM_ERR_INVALID;
    const unsigned nt = std::min(host_threads_available(), 16u);
    if (!h->pool || h->pool->n != nt) {
        delete h->pool;
        h->pool = new (std::nothrow) tmn::HostPool(nt);
        if (!h->pool) return tmn::fail(h, TM_ERR_NOMEM, "host worker pool%s%s");
    }
    const size_t words = (256u << 20) / sizeof(double);
    std::vector<double> src(words, 1.0), dst(words);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        const auto t0 = std::chrono::steady_clock::now();
        const double *sp = src.data();
        double *dp = dst.data();
        h->pool->run([=](unsigned t, unsigned n) {
            const size_t per = (words + n - 1) / n, lo = std::min(words, t * per), hi = std::min(words, lo + per);
            for (size_t i = lo; i < hi; ++i) store_f64(dp + i, sp[i]);
#if defined(__x86_64__)
            _mm_sfence();
#endif
        });
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        best = std::max(best, 2.0 * static_cast<double>(words) * sizeof(double) / s);
    }
    *bytes_per_second = best;
    if (threads) *threads = static_cast<int32_t>(nt);
    return TM_OK;
}

int tm_host_pipeline_info(tm_handle *h, int32_t *d2h_bytes_per_point, int32_t *host_threads) {
    if (!h) return TM_ERR_INVALID;
    if (d2h_bytes_per_point) *d2h_bytes_per_point = h->host_d2h_bytes_per_point;
    if (host_threads) *host_threads = h->host_assembly_threads;
    return TM_OK;
}

int tm_get_stats(tm_handle *h, tm_stats *out) {
    if (!h || !out) return TM_ERR_INVALID;
    TM_CUDA(h, cudaSetDevice(h->device));
    TM_CUDA(h, cudaDeviceSynchronize());
    tm_stats s = h->stats;
    if (s.mode_used == TM_MODE_GRID && h->dstats.p && h->n_listed + h->n_long > 0) {
        tmn::DevStats d;
        TM_CUDA(h, cudaMemcpy(&d, h->dstats.p, sizeof(d), cudaMemcpyDeviceToHost));
        s.pairs_evaluated += d.pairs_grid + d.pairs_ring;
        s.cull_tests += d.cull_tests;
        s.points_grid += static_cast<uint64_t>(h->last_n) - d.pending - d.far_certified;
        s.points_far += d.far_certified;
        s.points_ring += d.ring_certified;
        s.points_tree += d.pending - d.n_brute - d.ring_certified;
        s.points_brute += d.n_brute;
        s.index_entries = h->index_entries;
        s.voxels_occupied = d.voxels_occupied;
        s.work_items = d.work_items;
        s.bound_tests += d.bound_tests;
        s.points_slow += static_cast<uint64_t>(d.undecided_near) + d.undecided_far;
        s.lane_ops_per_bound = tmn::LANE_OPS_PER_BOUND;
        s.cell_size = h->grid.h;
        s.reach = h->reach;
        s.near_reach = h->near;
        s.grid_dim[0] = static_cast<uint32_t>(h->grid.nx);
        s.grid_dim[1] = static_cast<uint32_t>(h->grid.ny);
        s.grid_dim[2] = static_cast<uint32_t>(h->grid.nz);
    }
    *out = s;
    return TM_OK;
}

int tm_set_profiling(tm_handle *h, int enabled) {
    if (!h) return TM_ERR_INVALID;
    h->profiling = enabled != 0;
    return TM_OK;
}

int tm_get_phase_ms(tm_handle *h, float *out_ms) {
    if (!h || !out_ms) return TM_ERR_INVALID;
    for (int i = 0; i < TM_PHASES; ++i) out_ms[i] = 0.f;
    constexpr int END = TM_PHASES - 1;
    if (!h->profiling || !h->phase_hit[0] || !h->phase_hit[END]) return TM_OK;
    TM_CUDA(h, cudaSetDevice(h->device));
    TM_CUDA(h, cudaEventSynchronize(h->phase_ev[END]));
    // phase p spans mark p -> the next mark that was recorded (mark END = end of the call)
    for (int p = 0; p < END; ++p) {
        if (!h->phase_hit[p]) continue;
        int q = p + 1;
        while (q < END && !h->phase_hit[q]) ++q;
        if (p == 0 && !h->phase_hit[1]) continue;          // brute mode has no binning phase
        TM_CUDA(h, cudaEventElapsedTime(&out_ms[p], h->phase_ev[p], h->phase_ev[q]));
    }
    TM_CUDA(h, cudaEventElapsedTime(&out_ms[END], h->phase_ev[0], h->phase_ev[END]));
    return TM_OK;
}

int tm_selftest_arithmetic(tm_handle *h, uint64_t n, uint32_t seed, uint64_t *mismatches) {
    if (!h || !mismatches) return TM_ERR_INVALID;
    TM_CUDA(h, cudaSetDevice(h->device));
    TM_CUDA(h, h->scratch_f.ensure(256));
    unsigned long long *bad = h->scratch_f.as<unsigned long long>();
    TM_CUDA(h, cudaMemset(bad, 0, sizeof(unsigned long long)));
    selftest_kernel<<<h->sm_count * 8, 256>>>(n, seed, bad);
    TM_CUDA(h, cudaGetLastError());
    unsigned long long host = 0;
    TM_CUDA(h, cudaMemcpy(&host, bad, sizeof(host), cudaMemcpyDeviceToHost));
    *mismatches = host;
    return TM_OK;
}

int tm_measure_fp32_peak(tm_handle *h, double *lane_ops_per_second) {
    if (!h || !lane_ops_per_second) return TM_ERR_INVALID;
    TM_CUDA(h, cudaSetDevice(h->device));
    TM_CUDA(h, h->scratch_f.ensure(256));
    cudaEvent_t e0, e1;
    TM_CUDA(h, cudaEventCreate(&e0));
    TM_CUDA(h, cudaEventCreate(&e1));
    const int blocks = h->sm_count * 8, iters = 8192;
    double best = 0.0;
    for (int kind = 0; kind < 2; ++kind) {
        for (int rep = 0; rep < 4; ++rep) {
            TM_CUDA(h, cudaEventRecord(e0, 0));
            if (kind == 0) fp32_probe_kernel<0><<<blocks, 256>>>(h->scratch_f.as<float>(), iters, 1.0000001f, 1e-7f);
            else fp32_probe_kernel<1><<<blocks, 256>>>(h->scratch_f.as<float>(), iters, 1.0000001f, 1e-7f);
            TM_CUDA(h, cudaEventRecord(e1, 0));
            TM_CUDA(h, cudaEventSynchronize(e1));
            float ms = 0.f;
            TM_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
            const double ops = static_cast<double>(blocks) * 256.0 * iters * 8.0;
            if (rep > 0 && ms > 0.f) best = std::max(best, ops / (ms * 1e-3));
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    TM_CUDA(h, cudaGetLastError());
    *lane_ops_per_second = best;
    return TM_OK;
}

}  // extern "C"

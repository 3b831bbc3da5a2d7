// Noisy surface clouds on a QSM, generated on the device (PreProcessing/NoiseDataGeneration.py:60-102).
//
// The reference expands every cylinder into `count` points (np.repeat, :60), draws three N-long variate arrays from
// numpy's global generator (:64-68), builds (N,3,3) rotation matrices by fancy indexing (:99) and transforms: ~200 bytes
// of host-memory traffic per point for 24 bytes of result.  The per-cylinder quantities (counts, rotations; O(M), float64,
// bit-for-bit the reference's numpy expressions) stay on the host in NoiseDataGeneration.py of this package; this file is
// the O(N) part:
//
//   one thread per point; the owning cylinder is found by bisection over the exclusive prefix of the counts, narrowed to
//   the block's own range first; the variates are either read from caller-supplied arrays (how the parity tests replay the
//   reference's Mersenne-Twister draws) or drawn from Philox4x32-10 with the point index as the counter, so a cloud is a
//   pure function of (table, seed) however it is sharded; float64 arithmetic in the reference's order, no contraction;
//   rows leave through shared memory so that the (N,3) stores are contiguous.
//
// Bound: instruction issue (two Philox blocks, double-precision sincos, cos, log, exp, sqrt per point: 68 % of issue slots,
// FP64 pipe 27 %), far above its HBM writes of 24 B/point (+12 B for the optional float32 copy that feeds the labeller,
// Modules/Utils.py:236).
#include <cstdint>

#include "tm_core.cuh"

namespace tmn {

constexpr int NOISE_THREADS = 256;
constexpr int NOISE_REC = 14;                 // start xyz, rotation row-major (9), radius, length

struct Philox {
    uint32_t c[4];
};

__device__ __forceinline__ Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    Philox s{{c0, c1, c2, c3}};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, s.c[0]), lo0 = 0xD2511F53u * s.c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, s.c[2]), lo1 = 0xCD9E8D57u * s.c[2];
        const uint32_t n0 = hi1 ^ s.c[1] ^ k0, n2 = hi0 ^ s.c[3] ^ k1;
        s.c[0] = n0; s.c[1] = lo1; s.c[2] = n2; s.c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return s;
}

// two 32-bit words -> [0,1) with 53 random bits (the construction of numpy's legacy random_sample)
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return __ddiv_rn(__dadd_rn(__dmul_rn(static_cast<double>(a >> 5), 67108864.0), static_cast<double>(b >> 6)), 9007199254740992.0);
}

__global__ void __launch_bounds__(NOISE_THREADS)
noise_cloud_kernel(const double *__restrict__ rec, const int64_t *__restrict__ first, int64_t m, int64_t n, int64_t point0, uint32_t k0,
                   uint32_t k1, const double *__restrict__ theta_in, const double *__restrict__ z_in, const double *__restrict__ noise_in,
                   double *__restrict__ out, float *__restrict__ out32) {
    __shared__ double rows[NOISE_THREADS * 3];
    __shared__ int64_t range[2];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * NOISE_THREADS;
    const int64_t i = base + threadIdx.x;                       // row of this launch
    const int64_t last = min(base + NOISE_THREADS, n) - 1;
    if (threadIdx.x < 2) {
        // owner of the block's first / last point: largest c with first[c] <= g  (empty cylinders are skipped by the <=)
        const int64_t g = min(point0 + (threadIdx.x == 0 ? base : last), first[m] - 1);
        int64_t lo = 0, hi = m;                                 // first[lo] <= g < first[hi]
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (first[mid] <= g) lo = mid; else hi = mid;
        }
        range[threadIdx.x] = lo;
    }
    __syncthreads();
    if (i < n && point0 + i >= first[m]) {
        // a row past the end of the cloud (the caller asked for more rows than the plan holds): NaN, not another point
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        rows[threadIdx.x * 3] = rows[threadIdx.x * 3 + 1] = rows[threadIdx.x * 3 + 2] = nan;
    } else if (i < n) {
        const int64_t g = point0 + i;                           // global point number: the Philox counter
        int64_t lo = range[0], hi = range[1] + 1;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (first[mid] <= g) lo = mid; else hi = mid;
        }
        const double *c = rec + lo * NOISE_REC;
        const double radius = c[12], length = c[13];
        double theta, z, noise;
        if (theta_in) {
            theta = theta_in[i];
            z = z_in[i];
            noise = noise_in[i];
        } else {
            const uint32_t glo = static_cast<uint32_t>(g), ghi = static_cast<uint32_t>(static_cast<uint64_t>(g) >> 32);
            const Philox a = philox4x32_10(glo, ghi, 0u, 0u, k0, k1);
            const Philox b = philox4x32_10(glo, ghi, 1u, 0u, k0, k1);
            const double two_pi = 6.283185307179586;            // 2 * np.pi
            theta = __dmul_rn(two_pi, u53(a.c[0], a.c[1]));     // uniform(0, 2 pi)   (:64)
            z = __dmul_rn(length, u53(a.c[2], a.c[3]));         // uniform(0, L)      (:65)
            const double u1 = __dsub_rn(1.0, u53(b.c[0], b.c[1]));                          // (0, 1]
            const double gauss = __dmul_rn(sqrt(__dmul_rn(-2.0, log(u1))), cos(__dmul_rn(two_pi, u53(b.c[2], b.c[3]))));
            noise = exp(__dadd_rn(-3.0, __dmul_rn(0.85, gauss)));                          // lognormal(-3, 0.85)  (:68)
        }
        const double rho = __dadd_rn(radius, noise);            // :69
        double sn, cs;
        sincos(theta, &sn, &cs);
        const double x = __dmul_rn(rho, cs), y = __dmul_rn(rho, sn);                        // :72-73
#pragma unroll
        for (int r = 0; r < 3; ++r) {                           // R @ local + start  (:99, :102)
            const double dot = __dadd_rn(__dadd_rn(__dmul_rn(c[3 + 3 * r], x), __dmul_rn(c[4 + 3 * r], y)), __dmul_rn(c[5 + 3 * r], z));
            rows[threadIdx.x * 3 + r] = __dadd_rn(dot, c[r]);
        }
    }
    __syncthreads();
    const int64_t cnt = (last - base + 1) * 3;
    for (int64_t e = threadIdx.x; e < cnt; e += NOISE_THREADS) {
        const double v = rows[e];
        out[base * 3 + e] = v;
        if (out32) out32[base * 3 + e] = static_cast<float>(v);
    }
}

}  // namespace tmn

using namespace tmn;

extern "C" {

int tm_noise_cloud(tm_handle *h, const double *cyl_rec, const int64_t *first_point, int64_t m, int64_t n, int64_t point0, uint64_t seed,
                   const double *theta, const double *z, const double *noise, double *out_points, float *out_points_f32, void *stream) {
    if (!h) return TM_ERR_INVALID;
    if (n < 0 || m < 0 || point0 < 0) return fail(h, TM_ERR_INVALID, "tm_noise_cloud: negative size%s%s");
    if (n == 0) return TM_OK;
    if (m == 0) return fail(h, TM_ERR_NO_CYLINDERS, "%s%s", tm_status_string(TM_ERR_NO_CYLINDERS));
    if (!cyl_rec || !first_point || !out_points) return fail(h, TM_ERR_INVALID, "tm_noise_cloud: null pointer%s%s");
    const int given = (theta != nullptr) + (z != nullptr) + (noise != nullptr);
    if (given != 0 && given != 3) return fail(h, TM_ERR_INVALID, "tm_noise_cloud: pass all three variate arrays or none%s%s");
    TM_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t blocks = (n + NOISE_THREADS - 1) / NOISE_THREADS;
    if (blocks > 0x7fffffffLL) return fail(h, TM_ERR_INVALID, "tm_noise_cloud: more than 2^31-1 blocks of points in one call%s%s");
    noise_cloud_kernel<<<static_cast<unsigned>(blocks), NOISE_THREADS, 0, st>>>(cyl_rec, first_point, m, n, point0, static_cast<uint32_t>(seed),
                                                                                static_cast<uint32_t>(seed >> 32), theta, z, noise,
                                                                                out_points, out_points_f32);
    TM_KCHECK(h, st, "noise_cloud_kernel");
    return TM_OK;
}

}  // extern "C"

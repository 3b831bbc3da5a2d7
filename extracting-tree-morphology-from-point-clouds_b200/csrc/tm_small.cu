// Small-table fast path: the call pattern of cylinder_proximity_based_segmentation
// (Modules/Pipeline/QSMFittingDepthFirst.py:1006-1094).  That function runs thousands of times per tree, each time
// against the handful of cylinders fitted last (M ~ 1-10) and a few hundred to 1e5 points selected from the SAME
// cloud, and keeps a single bit per point: distance-to-closest-cylinder < eps (:1084).  The reference pays, per 1024
// points, five tiny H2D copies, ~70 ATen launches over (N_b, M, 3) temporaries and three synchronous D2H copies.
//
// Here the cloud is uploaded once (tm_cloud_upload_host) and stays resident; a call ships the selected row indices
// and the raw cylinders, ONE kernel prepares the cylinders in shared memory exactly as :1043-1045 does
// (axis = end - start, axis_length = ||axis||, axis_unit = axis / axis_length), evaluates every (point, cylinder)
// pair in the reference's operation order, takes torch.argmin's winner and writes flag / distance / row.
#include <algorithm>

#include "tm_core.cuh"
#include "tm_eval.cuh"

namespace tmn {

constexpr int SMALL_THREADS = 256;

template <bool GUARD, bool NFMA>
__global__ void __launch_bounds__(SMALL_THREADS) proximity_kernel(SmallArgs a) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    float4 *sA = reinterpret_cast<float4 *>(small_smem);
    float4 *sB = sA + a.m;
    for (int c = threadIdx.x; c < a.m; c += SMALL_THREADS) {
        const float *r = a.cyl + 7 * c;
        // QSMFittingDepthFirst.py:1043-1045 (same ops as Projection.py:126-132 when axis_eps > 0)
        const float ax = sub(r[3], r[0]), ay = sub(r[4], r[1]), az = sub(r[5], r[2]);
        const float len = norm3<NFMA>(ax, ay, az);
        const float dv = (a.axis_eps > 0.f && len < a.axis_eps) ? a.axis_eps : len;
        sA[c] = make_float4(r[0], r[1], r[2], len);
        sB[c] = make_float4(__fdiv_rn(ax, dv), __fdiv_rn(ay, dv), __fdiv_rn(az, dv), r[6]);
    }
    __syncthreads();
    for (int64_t i = blockIdx.x * static_cast<int64_t>(SMALL_THREADS) + threadIdx.x; i < a.n;
         i += static_cast<int64_t>(gridDim.x) * SMALL_THREADS) {
        const int64_t row = a.subset ? a.subset[i] : i;
        const float *p = a.cloud + 3 * row;
        const float px = p[0], py = p[1], pz = p[2];
        float bestd = __int_as_float(0x7f800000);
        int besti = 0;
        for (int j = 0; j < a.m; ++j) {
            const float d = eval_pair<GUARD, NFMA, false>(px, py, pz, sA[j], sB[j], a.atol, a.eps_norm, nullptr);
            // ascending index order: torch.argmin keeps the first NaN, else the first minimum
            const bool wins = !(d >= bestd) && (bestd == bestd);
            bestd = wins ? d : bestd;
            besti = wins ? j : besti;
        }
        if (a.flags) a.flags[i] = bestd < a.eps_flag ? 1 : 0;          // distances_batch < eps (:1084); NaN -> False
        if (a.dist) a.dist[i] = bestd;
        if (a.index) a.index[i] = besti;
    }
}

int run_proximity(tm_handle *h, const SmallArgs &a, bool guard, bool nfma, cudaStream_t st) {
    const size_t smem = sizeof(float4) * 2 * static_cast<size_t>(a.m);
    const int blocks = static_cast<int>(std::min<int64_t>((a.n + SMALL_THREADS - 1) / SMALL_THREADS, static_cast<int64_t>(h->sm_count) * 8));
#define TM_SMALL_CASE(G, F)                                                                                            \
    do {                                                                                                               \
        if (smem > 48 * 1024)                                                                                          \
            TM_CUDA(h, cudaFuncSetAttribute(proximity_kernel<G, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                            static_cast<int>(smem)));                                                  \
        proximity_kernel<G, F><<<blocks, SMALL_THREADS, smem, st>>>(a);                                                \
    } while (0)
    if (guard) { if (nfma) TM_SMALL_CASE(true, true); else TM_SMALL_CASE(true, false); }
    else       { if (nfma) TM_SMALL_CASE(false, true); else TM_SMALL_CASE(false, false); }
#undef TM_SMALL_CASE
    TM_KCHECK(h, st, "proximity_kernel");
    return TM_OK;
}

}  // namespace tmn

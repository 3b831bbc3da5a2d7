// Exhaustive nearest-cylinder kernel: every (point, cylinder) pair, cylinders streamed through
// shared memory in TMA-staged tiles.  It is
//   * the FP32-roofline yard-stick (pairs = N*M exactly, no accounting ambiguity),
//   * the small-M path of the QSM-fitting call site (QSMFittingDepthFirst.py:1079-1081),
//   * the fallback for points the voxel grid cannot answer (outside the grid, non-finite, voxels
//     whose candidate tile would be too large).
// Replaces the (N_b, M, 3) broadcast chain of LabelGenerationCuda.py:36-88 / Projection.py:35-91.
#include <algorithm>

#include "tm_core.cuh"
#include "tm_eval.cuh"
#include "tm_ptx.cuh"

namespace tmn {

constexpr int BRUTE_THREADS = 256;
constexpr int BRUTE_TILE = 512;          // cylinders per stage: 2 x 8 KB (A and B records)
constexpr int BRUTE_STAGES = 2;

// One thread owns P points (registers); the CTA walks the cylinder table tile by tile.  A single
// elected thread arms the stage's mbarrier and issues two bulk copies (A records, B records) for
// tile t+1 while all threads evaluate tile t from shared memory: every LDS is a 128-bit broadcast.
// gridDim.y splits the table so small point sets still fill the machine; partial winners meet in
// a 64-bit atomicMin on the (distance, index) key, which is torch.argmin's comparator.
template <int P, bool GUARD, bool NFMA>
__global__ void __launch_bounds__(BRUTE_THREADS)
brute_kernel(const float *__restrict__ pts, int64_t n, int64_t row_stride, const int32_t *__restrict__ sel,
             const unsigned int *__restrict__ d_count, const float4 *__restrict__ recA,
             const float4 *__restrict__ recB, int m, int tiles_per_split, float atol, float eps,
             unsigned long long *__restrict__ keys, int use_atomic) {
    __shared__ __align__(128) float4 sA[BRUTE_STAGES][BRUTE_TILE];
    __shared__ __align__(128) float4 sB[BRUTE_STAGES][BRUTE_TILE];
    __shared__ __align__(8) uint64_t full[BRUTE_STAGES];

    const int tid = threadIdx.x;
    const int64_t n_eff = d_count ? static_cast<int64_t>(*d_count) : n;
    const int ntiles = (m + BRUTE_TILE - 1) / BRUTE_TILE;
    const int t_begin = blockIdx.y * tiles_per_split;
    const int t_end = min(ntiles, t_begin + tiles_per_split);
    if (t_begin >= t_end) return;

    if (tid == 0) {
        for (int s = 0; s < BRUTE_STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t issued = 0, consumed = 0;       // running tile counters → stage = c % STAGES, parity = (c / STAGES) & 1
    auto issue = [&](int t) {
        const int s = issued % BRUTE_STAGES;
        const int base = t * BRUTE_TILE;
        const uint32_t cnt = static_cast<uint32_t>(min(BRUTE_TILE, m - base));
        mbar_expect_tx(&full[s], cnt * 32u);
        bulk_g2s(&sA[s][0], recA + base, cnt * 16u, &full[s]);
        bulk_g2s(&sB[s][0], recB + base, cnt * 16u, &full[s]);
    };

    constexpr int PTS = BRUTE_THREADS * P;
    for (int64_t blk = blockIdx.x; blk * PTS < n_eff; blk += gridDim.x) {
        float px[P], py[P], pz[P], bestd[P];
        uint32_t besti[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
            int64_t slot = blk * PTS + k * BRUTE_THREADS + tid;
            if (slot >= n_eff) slot = n_eff - 1;                    // duplicate work, masked at the store
            const int64_t row = sel ? static_cast<int64_t>(sel[slot]) : slot;
            const float *p = pts + row * row_stride;
            px[k] = p[0]; py[k] = p[1]; pz[k] = p[2];
            bestd[k] = __int_as_float(0x7f800000);                  // +inf
            besti[k] = static_cast<uint32_t>(t_begin) * BRUTE_TILE; // all-inf row → lowest index
        }

        if (tid == 0) { issue(t_begin); }
        ++issued;
        for (int t = t_begin; t < t_end; ++t) {
            if (t + 1 < t_end) {
                if (tid == 0) issue(t + 1);     // its buffer was released by the barrier ending tile t-1
                ++issued;
            }
            const int s = consumed % BRUTE_STAGES;
            mbar_wait(&full[s], (consumed / BRUTE_STAGES) & 1);
            ++consumed;
            const int base = t * BRUTE_TILE;
            const int cnt = min(BRUTE_TILE, m - base);
#pragma unroll 2
            for (int j = 0; j < cnt; ++j) {
                const float4 a = sA[s][j];
                const float4 b = sB[s][j];
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const float d = eval_pair<GUARD, NFMA, false>(px[k], py[k], pz[k], a, b, atol, eps, nullptr);
                    // ascending index order: the newcomer wins iff the incumbent is not NaN and
                    // (newcomer is NaN or strictly smaller)  — LessOrNan with lowest-index ties
                    const bool wins = !(d >= bestd[k]) && (bestd[k] == bestd[k]);
                    bestd[k] = wins ? d : bestd[k];
                    besti[k] = wins ? static_cast<uint32_t>(base + j) : besti[k];
                }
            }
            __syncthreads();                    // everyone is done with stage s → it may be refilled
            fence_proxy_async();
        }

#pragma unroll
        for (int k = 0; k < P; ++k) {
            const int64_t slot = blk * PTS + k * BRUTE_THREADS + tid;
            if (slot < n_eff) {
                const unsigned long long key = make_key(bestd[k], besti[k]);
                if (use_atomic) atomicMin(keys + slot, key);
                else keys[slot] = key;
            }
        }
    }
}

// Winner-only epilogue for the exhaustive path (fused label + offset write, A:92-109): recompute the
// winning pair with full geometry (bit-identical distance), move to the mantle, gather the ID.
template <bool GUARD, bool NFMA>
__global__ void __launch_bounds__(256)
finalize_kernel(const float *__restrict__ pts, int64_t n, int64_t row_stride, const int32_t *__restrict__ sel,
                const unsigned int *__restrict__ d_count, const unsigned long long *__restrict__ keys,
                const float4 *__restrict__ recA, const float4 *__restrict__ recB, const int32_t *__restrict__ ids,
                float atol, float eps, int move_to_mantle, int32_t *__restrict__ out_index,
                int32_t *__restrict__ out_id, float *__restrict__ out_dist, float *__restrict__ out_offset,
                float *__restrict__ out_radius, float4 *__restrict__ out_packed) {
    const int64_t n_eff = d_count ? static_cast<int64_t>(*d_count) : n;
    for (int64_t slot = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; slot < n_eff;
         slot += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = sel ? static_cast<int64_t>(sel[slot] & 0x3fffffff) : slot;   // sign bit: "outside the grid" flag
        const uint32_t j = key_index(keys[slot]);
        const float *p = pts + row * row_stride;
        const float px = p[0], py = p[1], pz = p[2];
        const float4 a = recA[j], b = recB[j];
        PairGeom g;
        eval_pair<GUARD, NFMA, true>(px, py, pz, a, b, atol, eps, &g);
        float ox, oy, oz;
        mantle_offset<NFMA>(g, px, py, pz, move_to_mantle != 0, ox, oy, oz);
        if (out_index) out_index[row] = static_cast<int32_t>(j);
        if (out_id) out_id[row] = ids[j];
        if (out_dist) out_dist[row] = g.dist;
        if (out_offset) { out_offset[3 * row] = ox; out_offset[3 * row + 1] = oy; out_offset[3 * row + 2] = oz; }
        if (out_radius) out_radius[row] = b.w;
        if (out_packed) out_packed[row] = make_float4(ox, oy, oz, __int_as_float(ids[j]));
    }
}

// Exhaustive search for the SHORT list of pending points nothing can bound (non-finite coordinates; the count lives on
// the device).
// One warp per (point, 1024-cylinder chunk) task: the lanes stride over the chunk's records straight from global
// memory (coalesced 512-byte requests, the table is L2 resident), skip candidates whose capsule lies beyond the
// incumbent (same cull as the tile kernel, with a per-point rounding allowance), keep a 64-bit (distance, index)
// key each, and a five-step warp-shuffle butterfly reduces them with torch.argmin's comparator (NaN first, then
// distance, then lowest index).  Chunks of one point meet in an atomicMin on the key.  The warp that owns chunk 0
// also evaluates the cylinders the cull cannot bound (special) and the axis-parallel ones (variant A).
constexpr int WARP_CHUNK = 1024;

struct BruteCullArgs {
    const float *pts;
    int64_t row_stride;
    const int32_t *pend_idx;
    const uint32_t *brute_slots;
    const unsigned int *d_count;
    const float4 *recA, *recB;
    int m;
    const int32_t *special, *aligned;
    uint32_t n_special, n_aligned;
    float atol, eps, maxabs, slack_floor;
    unsigned long long *keys;
    DevStats *st;
};

template <bool GUARD, bool NFMA>
__global__ void __launch_bounds__(256) brute_cull_kernel(BruteCullArgs a) {
    const unsigned int n_eff = *a.d_count;
    if (n_eff == 0) return;
    const int lane = threadIdx.x & 31;
    const unsigned long long nchunks = (static_cast<unsigned long long>(a.m) + WARP_CHUNK - 1) / WARP_CHUNK;
    const unsigned long long total = static_cast<unsigned long long>(n_eff) * nchunks;
    const unsigned long long nwarps = (static_cast<unsigned long long>(gridDim.x) * blockDim.x) >> 5;
    unsigned long long pairs = 0, culls = 0;
    for (unsigned long long task = (static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; task < total;
         task += nwarps) {
        const unsigned int slot = a.brute_slots[static_cast<unsigned int>(task / nchunks)];
        const int chunk = static_cast<int>(task % nchunks);
        const float *p = a.pts + static_cast<int64_t>(a.pend_idx[slot] & 0x3fffffff) * a.row_stride;
        const float px = p[0], py = p[1], pz = p[2];
        // rounding allowance at this point's own coordinate scale (inf / NaN coordinates: nothing is culled)
        const float slack = a.slack_floor + 4e-6f * fmaxf(fmaxf(fabsf(px), fabsf(py)), fmaxf(fabsf(pz), a.maxabs));
        const unsigned long long key0 = a.keys[slot];          // incumbent from the other chunks (may be stale)
        unsigned long long best = KEY_NONE;
        float thr = thr_of(key0, slack);
        if (!(fabsf(px) + fabsf(py) + fabsf(pz) < 3.0e38f)) thr = __int_as_float(0x7fc00000);
        const int j_end = min(a.m, (chunk + 1) * WARP_CHUNK);
        for (int j = chunk * WARP_CHUNK + lane; j < j_end; j += 32) {
            const float4 ca = a.recA[j], cb = a.recB[j];
            if (cull_pass(px, py, pz, ca, cb, thr)) {
                const float d = eval_pair<GUARD, NFMA, false>(px, py, pz, ca, cb, a.atol, a.eps, nullptr);
                const unsigned long long k = make_key(d, static_cast<uint32_t>(j));
                best = k < best ? k : best;
                thr = thr_of(best < key0 ? best : key0, slack);
                ++pairs;
            }
            ++culls;
        }
        if (chunk == 0) {
            for (uint32_t e = lane; e < a.n_special; e += 32) {
                const uint32_t j = static_cast<uint32_t>(a.special[e]);
                const unsigned long long k = make_key(eval_pair<GUARD, NFMA, false>(px, py, pz, a.recA[j], a.recB[j], a.atol, a.eps, nullptr), j);
                best = k < best ? k : best;
                ++pairs;
            }
            if (!GUARD) {
                for (uint32_t e = lane; e < a.n_aligned; e += 32) {
                    const uint32_t j = static_cast<uint32_t>(a.aligned[e]);
                    const float4 ca = a.recA[j], cb = a.recB[j];
                    if (on_axis_line(px, py, pz, ca, cb)) {
                        const unsigned long long k = make_key(eval_pair<GUARD, NFMA, false>(px, py, pz, ca, cb, a.atol, a.eps, nullptr), j);
                        best = k < best ? k : best;
                        ++pairs;
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other < best ? other : best;
        }
        if (lane == 0 && best != KEY_NONE) atomicMin(a.keys + slot, best);
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

template <int P, bool GUARD, bool NFMA>
static cudaError_t launch_brute(dim3 grid, cudaStream_t st, const float *pts, int64_t n, int64_t rs, const int32_t *sel,
                                const unsigned int *d_count, const float4 *A, const float4 *B, int m, int tps, float atol,
                                float eps, unsigned long long *keys, int use_atomic) {
    brute_kernel<P, GUARD, NFMA><<<grid, BRUTE_THREADS, 0, st>>>(pts, n, rs, sel, d_count, A, B, m, tps, atol, eps, keys,
                                                                use_atomic);
    return cudaGetLastError();
}

static int run_brute(tm_handle *h, const LabelArgs &a, const int32_t *sel, const unsigned int *d_count, int64_t n_launch) {
    const bool guard = a.prm.norm_eps > 0.f;
    const bool nfma = a.prm.norm_fma != 0;
    const int m = static_cast<int>(h->m);
    const int ntiles = (m + BRUTE_TILE - 1) / BRUTE_TILE;
    TM_CUDA(h, h->keys.ensure(sizeof(unsigned long long) * static_cast<size_t>(n_launch)));
    unsigned long long *keys = h->keys.as<unsigned long long>();

    int P, splits;
    int64_t nblk;
    {
        // points per thread: 2 once there are enough points to fill the machine (the two 128-bit broadcast loads of a
        // cylinder then serve two evaluations: 0.60 -> 0.69 of the FP32 issue roofline, profiles/r02_brute_kernel.md;
        // 4 points per thread measured within noise of 2: TM_BRUTE_P selects it for experiments)
        const int64_t full_wave = static_cast<int64_t>(h->sm_count) * 3 * BRUTE_THREADS;
        P = (n_launch >= 2 * full_wave) ? 2 : 1;
        if (const char *env = getenv("TM_BRUTE_P")) { const int v = atoi(env); if (v == 1 || v == 2 || v == 4) P = v; }
        nblk = (n_launch + BRUTE_THREADS * P - 1) / (BRUTE_THREADS * P);
        const int64_t want_ctas = static_cast<int64_t>(h->sm_count) * 6;
        splits = 1;
        if (nblk < want_ctas) splits = static_cast<int>(std::min<int64_t>(ntiles, (want_ctas + nblk - 1) / nblk));
    }
    if (splits > 65535) splits = 65535;
    const int tps = (ntiles + splits - 1) / splits;
    splits = (ntiles + tps - 1) / tps;
    const int use_atomic = splits > 1;
    if (use_atomic) TM_CUDA(h, cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * static_cast<size_t>(n_launch), a.stream));
    dim3 grid(static_cast<unsigned>(std::min<int64_t>(nblk, 1 << 20)), static_cast<unsigned>(splits));
    const float4 *A = h->recA.as<float4>();
    const float4 *B = h->recB.as<float4>();
    if (!d_count) mark(h, 5, a.stream);
    cudaError_t e;
#define TM_BRUTE_CASE(PP, G, F)                                                                                   \
    e = launch_brute<PP, G, F>(grid, a.stream, a.pts, n_launch, a.row_stride, sel, d_count, A, B, m, tps,           \
                               a.prm.perp_atol, a.prm.norm_eps, keys, use_atomic)
    if (P == 4) {
        if (guard) { if (nfma) TM_BRUTE_CASE(4, true, true); else TM_BRUTE_CASE(4, true, false); }
        else       { if (nfma) TM_BRUTE_CASE(4, false, true); else TM_BRUTE_CASE(4, false, false); }
    } else if (P == 2) {
        if (guard) { if (nfma) TM_BRUTE_CASE(2, true, true); else TM_BRUTE_CASE(2, true, false); }
        else       { if (nfma) TM_BRUTE_CASE(2, false, true); else TM_BRUTE_CASE(2, false, false); }
    } else {
        if (guard) { if (nfma) TM_BRUTE_CASE(1, true, true); else TM_BRUTE_CASE(1, true, false); }
        else       { if (nfma) TM_BRUTE_CASE(1, false, true); else TM_BRUTE_CASE(1, false, false); }
    }
#undef TM_BRUTE_CASE
    TM_CUDA(h, e);

    mark(h, 6, a.stream);
    const int fgrid = static_cast<int>(std::min<int64_t>((n_launch + 255) / 256, static_cast<int64_t>(h->sm_count) * 16));
#define TM_FIN_CASE(G, F)                                                                                          \
    finalize_kernel<G, F><<<fgrid, 256, 0, a.stream>>>(a.pts, n_launch, a.row_stride, sel, d_count, keys, A, B,     \
                                                       h->ids.as<int32_t>(), a.prm.perp_atol, a.prm.norm_eps,       \
                                                       a.prm.move_to_mantle, a.out_index, a.out_id, a.out_dist,     \
                                                       a.out_offset, a.out_radius, a.out_packed)
    if (guard) { if (nfma) TM_FIN_CASE(true, true); else TM_FIN_CASE(true, false); }
    else       { if (nfma) TM_FIN_CASE(false, true); else TM_FIN_CASE(false, false); }
#undef TM_FIN_CASE
    TM_CUDA(h, cudaGetLastError());
    h->stats.launches += 2;                  // exhaustive kernel + its winner epilogue
    return TM_OK;
}

int label_brute(tm_handle *h, const LabelArgs &a) {
    if (a.n == 0) return TM_OK;
    int rc = run_brute(h, a, nullptr, nullptr, a.n);
    if (rc != TM_OK) return rc;
    h->stats.pairs_evaluated += static_cast<uint64_t>(a.n) * static_cast<uint64_t>(h->m);
    h->stats.points_brute += static_cast<uint64_t>(a.n);
    return TM_OK;
}

// winning rows of the pending slots the exhaustive kernel settled (non-finite points) -> win[original row]; the ring and
// tree searches write theirs themselves
__global__ void __launch_bounds__(256) pending_winner_kernel(const int32_t *__restrict__ pend_idx, const uint32_t *__restrict__ brute_slots,
                                                             const unsigned int *__restrict__ d_count,
                                                             const unsigned long long *__restrict__ keys, int32_t *__restrict__ win) {
    const unsigned int n = *d_count;
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const unsigned int slot = brute_slots[k];
        win[pend_idx[slot] & 0x3fffffff] = static_cast<int32_t>(key_index(keys[slot]));
    }
}

// Streaming winner-only epilogue (A:92-109) over the rows in input order: recompute the winning pair with full geometry
// (bit-identical distance), move to the mantle, gather the ID.  Reads 12 + 4 bytes per point, writes the outputs; the
// 16-byte cylinder records come from L1/L2 (neighbouring rows of a sorted scan line mostly share their winner, random
// rows still hit the 1.6 MB table in L2).
template <bool GUARD, bool NFMA>
__global__ void __launch_bounds__(256)
finalize_rows_kernel(const float *__restrict__ pts, int64_t n, int64_t row_stride, const int32_t *__restrict__ win,
                     const float4 *__restrict__ recAB, const int32_t *__restrict__ ids,
                     float atol, float eps, int move_to_mantle, int32_t *__restrict__ out_index, int32_t *__restrict__ out_id,
                     float *__restrict__ out_dist, float *__restrict__ out_offset, float *__restrict__ out_radius,
                     float4 *__restrict__ out_packed, const uint32_t *__restrict__ rows, const unsigned int *__restrict__ d_count) {
    // two rows per thread and iteration: both rows' loads (point, winning row, then the dependent record gathers) are in
    // flight together, which is what hides the L2 latency of the gathers
    constexpr int R = 2;
    const int64_t span = static_cast<int64_t>(gridDim.x) * blockDim.x;
    if (rows) n = *d_count;                       // list mode: the rows to finish are rows[0 .. *d_count)
    for (int64_t row0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; row0 < n; row0 += R * span) {
        int64_t row[R];
        bool ok[R];
        uint32_t j[R];
        float px[R], py[R], pz[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            row[k] = row0 + k * span;
            ok[k] = row[k] < n;
            if (rows) row[k] = rows[ok[k] ? row[k] : row0] & 0x3fffffffu;        // pending slots carry two flag bits
            const int64_t r = (ok[k] || rows) ? row[k] : row0;
            j[k] = static_cast<uint32_t>(win[r]);
            const float *p = pts + r * row_stride;
            px[k] = p[0]; py[k] = p[1]; pz[k] = p[2];
        }
        float4 a[R], b[R];
        int32_t id[R];
#pragma unroll
        for (int k = 0; k < R; ++k) { a[k] = recAB[2 * j[k]]; b[k] = recAB[2 * j[k] + 1]; id[k] = (out_id || out_packed) ? ids[j[k]] : 0; }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            PairGeom g;
            eval_pair<GUARD, NFMA, true>(px[k], py[k], pz[k], a[k], b[k], atol, eps, &g);
            float ox, oy, oz;
            mantle_offset<NFMA>(g, px[k], py[k], pz[k], move_to_mantle != 0, ox, oy, oz);
            if (!ok[k]) continue;
            const int64_t r = row[k];
            if (out_index && out_index != win) out_index[r] = static_cast<int32_t>(j[k]);
            if (out_id) out_id[r] = id[k];
            if (out_dist) out_dist[r] = g.dist;
            if (out_offset) { out_offset[3 * r] = ox; out_offset[3 * r + 1] = oy; out_offset[3 * r + 2] = oz; }
            if (out_radius) out_radius[r] = b[k].w;
            if (out_packed) out_packed[r] = make_float4(ox, oy, oz, __int_as_float(id[k]));
        }
    }
}

int finalize_rows(tm_handle *h, const LabelArgs &a, const int32_t *win, const uint32_t *rows, const unsigned int *d_count) {
    const bool guard = a.prm.norm_eps > 0.f, nfma = a.prm.norm_fma != 0;
    const int grid = static_cast<int>(std::min<int64_t>((a.n + 511) / 512, static_cast<int64_t>(h->sm_count) * 16));
#define TM_FINR_CASE(G, F)                                                                                          \
    finalize_rows_kernel<G, F><<<grid, 256, 0, a.stream>>>(a.pts, a.n, a.row_stride, win, h->recAB.as<float4>(),     \
                                                           h->ids.as<int32_t>(), a.prm.perp_atol,                       \
                                                           a.prm.norm_eps, a.prm.move_to_mantle, a.out_index, a.out_id, \
                                                           a.out_dist, a.out_offset, a.out_radius, a.out_packed, rows, d_count)
    if (guard) { if (nfma) TM_FINR_CASE(true, true); else TM_FINR_CASE(true, false); }
    else       { if (nfma) TM_FINR_CASE(false, true); else TM_FINR_CASE(false, false); }
#undef TM_FINR_CASE
    TM_KCHECK(h, a.stream, "finalize_rows_kernel");
    h->stats.launches += 1;
    return TM_OK;
}

int finish_pending(tm_handle *h, const LabelArgs &a, DevStats *dst, int32_t *win, float maxabs) {
    const bool guard = a.prm.norm_eps > 0.f, nfma = a.prm.norm_fma != 0;
    BruteCullArgs b;
    b.pts = a.pts; b.row_stride = a.row_stride;
    b.pend_idx = h->pend_idx.as<int32_t>();
    b.brute_slots = h->brute_slots.as<uint32_t>();
    b.d_count = &dst->n_brute;
    b.recA = h->recA.as<float4>(); b.recB = h->recB.as<float4>();
    b.m = static_cast<int>(h->m);
    b.special = h->special.as<int32_t>(); b.aligned = h->aligned.as<int32_t>();
    b.n_special = h->n_special; b.n_aligned = h->n_aligned;
    b.atol = a.prm.perp_atol; b.eps = a.prm.norm_eps; b.maxabs = maxabs; b.slack_floor = h->slack_floor;
    b.keys = h->keys.as<unsigned long long>();
    b.st = dst;
    // non-finite points are rare: a small grid, and grid-stride loops if there are many after all
    const int wgrid = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(h->sm_count) * 8, std::max<int64_t>(h->sm_count / 2, a.n / 4096)));
    if (guard) { if (nfma) brute_cull_kernel<true, true><<<wgrid, 256, 0, a.stream>>>(b); else brute_cull_kernel<true, false><<<wgrid, 256, 0, a.stream>>>(b); }
    else       { if (nfma) brute_cull_kernel<false, true><<<wgrid, 256, 0, a.stream>>>(b); else brute_cull_kernel<false, false><<<wgrid, 256, 0, a.stream>>>(b); }
    TM_KCHECK(h, a.stream, "brute_cull_kernel");
    mark(h, 6, a.stream);
    pending_winner_kernel<<<std::max(1, wgrid / 8), 256, 0, a.stream>>>(h->pend_idx.as<int32_t>(), h->brute_slots.as<uint32_t>(), &dst->n_brute,
                                                                       h->keys.as<unsigned long long>(), win);
    TM_KCHECK(h, a.stream, "pending_winner_kernel");
    h->stats.launches += 2;
    return TM_OK;
}

}  // namespace tmn

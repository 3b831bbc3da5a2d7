// Bounding-volume hierarchy over the cylinders' capsule boxes: the search structure for the points the voxel tiles
// cannot certify (clutter, foliage, ground: anything farther than D_max from every cylinder, or outside the grid).
//
// The tiles answer a surface-sampled point with ~17 culls; a point d metres away from the tree would have to look at
// every cylinder within d of its voxel — thousands in a crown — whereas a per-point tree descent with the incumbent as
// the pruning radius visits O(log M) boxes plus the few leaves that are nearly as close as the winner.
//
//   build (per table, device): Morton-order the regular cylinders by box centre (radix sort), split the order in the
//                             middle recursively, leaves of <= 4 cylinders, node = the two child boxes + child codes
//                             in one 64-byte record; boxes bottom-up, one kernel.
//   search (per call):        one THREAD per pending point, explicit stack in local memory, nearer child first, node
//                             pruned when dist(p, box) > thr(incumbent); leaf cylinders go through the same capsule cull
//                             and the reference-order evaluation as everywhere else.  The incumbent left by the tile
//                             kernel seeds the search.
//
// Exactness: capsule(c) lies inside box(c) (padded, tm_api.cu pack_kernel) and box(c) inside every ancestor's box, so
// dist_ref(p,c) >= dist(p, capsule(c)) >= dist(p, box(node)): a pruned subtree cannot beat or tie the incumbent.  The
// search is exhaustive up to pruning, so every point it handles is final.  Cylinders that cannot be bounded (special)
// and variant A's axis-parallel ones are not in the tree: the tile kernel has already evaluated them for the points it
// saw; for points outside the grid they are evaluated here.
#include <algorithm>
#include <cmath>

#include <cub/device/device_radix_sort.cuh>

#include "tm_core.cuh"
#include "tm_eval.cuh"

namespace tmn {

constexpr int BVH_LEAF = 4;
constexpr int BVH_STACK = 40;

struct __align__(16) BvhNode {
    float lo0[3], hi0[3];
    float lo1[3], hi1[3];
    int32_t c0, c1;          // >= 0: internal node; < 0: leaf, -1 - ((first << 3) | (count - 1))
    int32_t pad[2];
};
static_assert(sizeof(BvhNode) == 64, "one node = two 32-byte sectors");

// ---- device build ---------------------------------------------------------------------------------------------
// The tree over n Morton-ordered cylinders is the implicit "split the range in the middle" tree: its shape depends on n
// alone, so nodes are numbered like a binary heap (root 1, children 2i and 2i+1) and every thread finds the range of its
// slot by following the bits of the slot number from the root.  Boxes come bottom-up: one thread per leaf computes the
// leaf's box and climbs; at every parent the second arrival (atomic counter) has both children's boxes in front of it,
// writes the 64-byte node record and goes on.
__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

struct __align__(16) BvhSlot {       // build scratch per heap slot: the subtree's box and range
    float lo[3], hi[3];
    int32_t first, count;
};

// 64-bit sort key: Morton code of the box centre | row; special cylinders (not in the tree) sort to the end
__global__ void bvh_key_kernel(const float4 *__restrict__ boxlo, const float4 *__restrict__ boxhi, int m, float ox, float oy, float oz,
                               float sx, float sy, float sz, unsigned long long *__restrict__ keys) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const float4 a = boxlo[c], b = boxhi[c];
    if (!(a.w == 0.f)) { keys[c] = 0xFFFFFFFF00000000ull | static_cast<uint32_t>(c); return; }
    const uint32_t qx = static_cast<uint32_t>(fminf(fmaxf((0.5f * (a.x + b.x) - ox) * sx, 0.f), 1023.f));
    const uint32_t qy = static_cast<uint32_t>(fminf(fmaxf((0.5f * (a.y + b.y) - oy) * sy, 0.f), 1023.f));
    const uint32_t qz = static_cast<uint32_t>(fminf(fmaxf((0.5f * (a.z + b.z) - oz) * sz, 0.f), 1023.f));
    const uint32_t code = expand10(qx) | (expand10(qy) << 1) | (expand10(qz) << 2);
    keys[c] = (static_cast<unsigned long long>(code) << 32) | static_cast<uint32_t>(c);
}

__device__ __forceinline__ int32_t bvh_child_ref(int slot, const BvhSlot &s) {
    return s.count <= BVH_LEAF ? -1 - ((s.first << 3) | (s.count - 1)) : slot;
}

__global__ void bvh_build_kernel(const unsigned long long *__restrict__ sorted_keys, const float4 *__restrict__ boxlo,
                                 const float4 *__restrict__ boxhi, int n, int slots, int32_t *__restrict__ rows,
                                 BvhSlot *__restrict__ scratch, int *__restrict__ arrivals, BvhNode *__restrict__ nodes) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) rows[t] = static_cast<int32_t>(static_cast<uint32_t>(sorted_keys[t]));
    int id = t + 1;                                             // heap slot 1 .. slots-1
    if (id >= slots) return;
    // range of this slot: follow the bits of the slot number below its leading one
    int first = 0, count = n;
    const int depth = 31 - __clz(id);
    for (int b = depth - 1; b >= 0; --b) {
        if (count <= BVH_LEAF) return;                          // an ancestor is a leaf: the slot does not exist
        const int half = count / 2;
        if ((id >> b) & 1) { first += half; count -= half; } else { count = half; }
    }
    if (count > BVH_LEAF) return;                               // inner node: written by whichever child arrives second
    BvhSlot me;
    me.first = first; me.count = count;
    for (int k = 0; k < 3; ++k) { me.lo[k] = INFINITY; me.hi[k] = -INFINITY; }
    for (int i = first; i < first + count; ++i) {
        const uint32_t r = static_cast<uint32_t>(sorted_keys[i]);
        const float4 a = boxlo[r], b = boxhi[r];
        me.lo[0] = fminf(me.lo[0], a.x); me.lo[1] = fminf(me.lo[1], a.y); me.lo[2] = fminf(me.lo[2], a.z);
        me.hi[0] = fmaxf(me.hi[0], b.x); me.hi[1] = fmaxf(me.hi[1], b.y); me.hi[2] = fmaxf(me.hi[2], b.z);
    }
    for (;;) {
        scratch[id] = me;
        if (id == 1) return;                                    // the root's own box is not needed
        __threadfence();
        const int parent = id >> 1;
        if (atomicAdd(&arrivals[parent], 1) == 0) return;       // the sibling will finish the parent
        __threadfence();
        const volatile BvhSlot *vs = scratch;
        BvhSlot c0, c1;
        for (int k = 0; k < 3; ++k) {
            c0.lo[k] = vs[2 * parent].lo[k]; c0.hi[k] = vs[2 * parent].hi[k];
            c1.lo[k] = vs[2 * parent + 1].lo[k]; c1.hi[k] = vs[2 * parent + 1].hi[k];
        }
        c0.first = vs[2 * parent].first; c0.count = vs[2 * parent].count;
        c1.first = vs[2 * parent + 1].first; c1.count = vs[2 * parent + 1].count;
        BvhNode nd;
        for (int k = 0; k < 3; ++k) {
            nd.lo0[k] = c0.lo[k]; nd.hi0[k] = c0.hi[k]; nd.lo1[k] = c1.lo[k]; nd.hi1[k] = c1.hi[k];
            me.lo[k] = fminf(c0.lo[k], c1.lo[k]);
            me.hi[k] = fmaxf(c0.hi[k], c1.hi[k]);
        }
        nd.c0 = bvh_child_ref(2 * parent, c0);
        nd.c1 = bvh_child_ref(2 * parent + 1, c1);
        nd.pad[0] = nd.pad[1] = 0;
        nodes[parent] = nd;
        me.first = c0.first;
        me.count = c0.count + c1.count;
        id = parent;
    }
}

__global__ void bvh_leaf_gather_kernel(const int32_t *__restrict__ rows, int n, const float4 *__restrict__ recAB,
                                       float4 *__restrict__ leafAB) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = rows[i];
    leafAB[2 * i] = recAB[2 * r];
    leafAB[2 * i + 1] = recAB[2 * r + 1];
}

// n_regular cylinders (boxlo.w == 0) inside [lo, hi] (both known to the caller from the packing pass)
int build_bvh(tm_handle *h, cudaStream_t stream, int n_regular, const float *lo, const float *hi) {
    const int m = static_cast<int>(h->m);
    const int n = n_regular;
    h->bvh_count = 0;
    h->bvh_root = 0;
    if (m > (1 << 27)) return fail(h, TM_ERR_INVALID, "tm_set_cylinders: more than 2^27 cylinders%s%s");
    if (n <= 0) return TM_OK;
    int depth = 0;                                              // leaves live at depth <= `depth`
    while (((n + (1 << depth) - 1) >> depth) > BVH_LEAF) ++depth;
    const int slots = 2 << depth;                               // heap slots 1 .. 2^(depth+1) - 1
    float scale[3];
    for (int k = 0; k < 3; ++k) scale[k] = hi[k] > lo[k] ? 1023.999f / (hi[k] - lo[k]) : 0.f;

    size_t tmp_bytes = 0;
    unsigned long long *nokeys = nullptr;
    TM_CUDA(h, cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, nokeys, nokeys, m, 32, 64, stream));
    const size_t key_bytes = (sizeof(unsigned long long) * static_cast<size_t>(m) + 255) & ~static_cast<size_t>(255);
    const size_t slot_bytes = (sizeof(BvhSlot) * static_cast<size_t>(slots) + 255) & ~static_cast<size_t>(255);
    const size_t arr_bytes = (sizeof(int) * static_cast<size_t>(slots) + 255) & ~static_cast<size_t>(255);
    TM_CUDA(h, h->bvh_scratch.ensure(2 * key_bytes + slot_bytes + arr_bytes + tmp_bytes));
    unsigned char *base = h->bvh_scratch.as<unsigned char>();
    unsigned long long *keys_in = reinterpret_cast<unsigned long long *>(base);
    unsigned long long *keys_out = reinterpret_cast<unsigned long long *>(base + key_bytes);
    BvhSlot *scratch = reinterpret_cast<BvhSlot *>(base + 2 * key_bytes);
    int *arrivals = reinterpret_cast<int *>(base + 2 * key_bytes + slot_bytes);
    void *cub_tmp = base + 2 * key_bytes + slot_bytes + arr_bytes;
    TM_CUDA(h, h->bvh_nodes.ensure(sizeof(BvhNode) * static_cast<size_t>(slots)));
    TM_CUDA(h, h->bvh_rows.ensure(sizeof(int32_t) * static_cast<size_t>(n)));
    TM_CUDA(h, h->bvh_leafAB.ensure(sizeof(float4) * 2 * static_cast<size_t>(n)));

    bvh_key_kernel<<<(m + 255) / 256, 256, 0, stream>>>(h->boxlo.as<float4>(), h->boxhi.as<float4>(), m, lo[0], lo[1], lo[2], scale[0],
                                                       scale[1], scale[2], keys_in);
    TM_KCHECK(h, stream, "bvh_key_kernel");
    // stable sort on the code bits only: rows stay ascending among equal codes, special cylinders (all ones) end up last
    TM_CUDA(h, cub::DeviceRadixSort::SortKeys(cub_tmp, tmp_bytes, keys_in, keys_out, m, 32, 64, stream));
    TM_CUDA(h, cudaMemsetAsync(arrivals, 0, sizeof(int) * static_cast<size_t>(slots), stream));
    const int threads = std::max(n, slots);
    bvh_build_kernel<<<(threads + 255) / 256, 256, 0, stream>>>(keys_out, h->boxlo.as<float4>(), h->boxhi.as<float4>(), n, slots,
                                                                h->bvh_rows.as<int32_t>(), scratch, arrivals, h->bvh_nodes.as<BvhNode>());
    TM_KCHECK(h, stream, "bvh_build_kernel");
    bvh_leaf_gather_kernel<<<(n + 255) / 256, 256, 0, stream>>>(h->bvh_rows.as<int32_t>(), n, h->recAB.as<float4>(),
                                                              h->bvh_leafAB.as<float4>());
    TM_KCHECK(h, stream, "bvh_leaf_gather_kernel");
    h->bvh_root = n <= BVH_LEAF ? -1 - ((0 << 3) | (n - 1)) : 1;
    h->bvh_count = n;
    return TM_OK;
}

// ---- search ------------------------------------------------------------------------------------------------------
struct BvhArgs {
    const float *pts;
    int64_t row_stride;
    const int32_t *pend_idx;
    unsigned long long *pend_keys;
    const unsigned int *d_pending;
    int32_t *win;                  // winning row of every slot this kernel settles, at the point's row
    const BvhNode *nodes;
    const float4 *leafAB;
    const int32_t *leaf_rows;
    int32_t root;
    int32_t count;
    const float4 *recA, *recB;
    const int32_t *special, *aligned;
    uint32_t n_special, n_aligned;
    float atol, eps, maxabs, slack_floor;
    DevStats *st;
};

__device__ __forceinline__ float box_dist2(float px, float py, float pz, const float *lo, const float *hi) {
    const float gx = fmaxf(fmaxf(lo[0] - px, px - hi[0]), 0.f);
    const float gy = fmaxf(fmaxf(lo[1] - py, py - hi[1]), 0.f);
    const float gz = fmaxf(fmaxf(lo[2] - pz, pz - hi[2]), 0.f);
    return fmaf(gz, gz, fmaf(gy, gy, gx * gx));
}

// One point per LANE, but the lanes of a warp are kept together instead of each running its own loop:
//   * a lane whose search has ended takes the next pending slot right away (one atomic per warp and refill), so a warp
//     never waits for its slowest point;
//   * each round first lets the lanes that stand at inner nodes descend (a few steps, until every lane has reached a leaf
//     or run out of work) and then lets the lanes that stand at leaves run the cull + reference-order evaluation
//     together.  In a single "node or leaf" loop the expensive leaf body ran in almost every iteration for two or three
//     lanes (5.6 of 32 lanes active on average, profiles/r01i_tree_search.md).
constexpr int BVH_NODE_STEPS = 6;
constexpr int32_t BVH_IDLE = 0x7fffffff;
constexpr int32_t BVH_DONE_BIT = 0x40000000;      // tm_grid.cu: DONE_BIT of a pending slot

template <bool GUARD, bool NFMA>
__global__ void __launch_bounds__(128, 12) bvh_kernel(BvhArgs a) {
    const unsigned int n_pend = *a.d_pending;
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    unsigned long long pairs = 0, culls = 0;
    int32_t stk_node[BVH_STACK];
    float stk_d2[BVH_STACK];
    int sp = 0;
    int32_t node = BVH_IDLE;            // BVH_IDLE: no point in flight; >= 0: inner node to visit; < 0: leaf to visit
    unsigned int slot = 0;
    int32_t row = 0;
    float px = 0.f, py = 0.f, pz = 0.f, slack = 0.f, thr = 0.f;
    unsigned long long key = KEY_NONE;
    bool more = n_pend > 0;             // warp-uniform: slots left to hand out

    // next subtree that can still hold a winner, or the end of this point's search
    auto pop = [&]() {
        while (sp > 0) {
            --sp;
            if (!(stk_d2[sp] > thr * thr)) { node = stk_node[sp]; return; }
        }
        a.pend_keys[slot] = key;
        a.win[row] = static_cast<int32_t>(key_index(key));
        node = BVH_IDLE;
    };

    for (;;) {
        // ---- refill the idle lanes
        if (more) {
            const uint32_t want = __ballot_sync(0xffffffffu, node == BVH_IDLE);
            if (want) {
                const int leader = __ffs(want) - 1;
                unsigned int base = 0;
                if (lane == leader) base = atomicAdd(&a.st->bvh_cursor, static_cast<unsigned int>(__popc(want)));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (base + static_cast<unsigned int>(__popc(want)) >= n_pend) more = false;
                const unsigned int s = base + static_cast<unsigned int>(__popc(want & lt));
                const int32_t ri = (node == BVH_IDLE && s < n_pend) ? a.pend_idx[s] : BVH_DONE_BIT;
                if (!(ri & BVH_DONE_BIT)) {
                    const int32_t my_row = ri & 0x3fffffff;
                    const float *p = a.pts + static_cast<int64_t>(my_row) * a.row_stride;
                    px = p[0]; py = p[1]; pz = p[2];
                    key = a.pend_keys[s];
                    // non-finite points are the exhaustive kernel's; a NaN incumbent is final
                    if (fabsf(px) + fabsf(py) + fabsf(pz) < 3.0e38f && static_cast<uint32_t>(key >> 32) != 0u) {
                        slot = s;
                        row = my_row;
                        slack = a.slack_floor + 4e-6f * fmaxf(fmaxf(fabsf(px), fabsf(py)), fmaxf(fabsf(pz), a.maxabs));
                        if (ri < 0) {
                            // outside the grid: the tile kernel never saw this point
                            for (uint32_t e = 0; e < a.n_special; ++e) {
                                const uint32_t j = static_cast<uint32_t>(a.special[e]);
                                const unsigned long long k = make_key(eval_pair<GUARD, NFMA, false>(px, py, pz, a.recA[j], a.recB[j], a.atol, a.eps, nullptr), j);
                                key = k < key ? k : key;
                                ++pairs;
                            }
                            if (!GUARD) {
                                for (uint32_t e = 0; e < a.n_aligned; ++e) {
                                    const uint32_t j = static_cast<uint32_t>(a.aligned[e]);
                                    const float4 ca = a.recA[j], cb = a.recB[j];
                                    if (on_axis_line(px, py, pz, ca, cb)) {
                                        const unsigned long long k = make_key(eval_pair<GUARD, NFMA, false>(px, py, pz, ca, cb, a.atol, a.eps, nullptr), j);
                                        key = k < key ? k : key;
                                        ++pairs;
                                    }
                                }
                            }
                        }
                        if (a.count > 0 && static_cast<uint32_t>(key >> 32) != 0u) {
                            thr = thr_of(key, slack);           // NaN while there is no incumbent: nothing is pruned
                            sp = 0;
                            node = a.root;
                        } else {
                            a.pend_keys[s] = key;
                            a.win[my_row] = static_cast<int32_t>(key_index(key));
                        }
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, node != BVH_IDLE)) {
            if (!more) break;
            continue;
        }
        // ---- inner nodes: nearer child first, the other one on the stack
        for (int step = 0; step < BVH_NODE_STEPS; ++step) {
            const bool inner = node >= 0 && node != BVH_IDLE;
            if (!__any_sync(0xffffffffu, inner)) break;
            if (inner) {
                const float4 *q = reinterpret_cast<const float4 *>(a.nodes + node);
                const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
                const float lo0[3] = {q0.x, q0.y, q0.z}, hi0[3] = {q0.w, q1.x, q1.y};
                const float lo1[3] = {q1.z, q1.w, q2.x}, hi1[3] = {q2.y, q2.z, q2.w};
                const int32_t c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
                const float d0 = box_dist2(px, py, pz, lo0, hi0), d1 = box_dist2(px, py, pz, lo1, hi1);
                const float t2 = thr * thr;
                const bool h0 = !(d0 > t2), h1 = !(d1 > t2);
                if (h0 && h1) {
                    const bool first0 = d0 <= d1;
                    if (sp < BVH_STACK) { stk_node[sp] = first0 ? c1 : c0; stk_d2[sp] = first0 ? d1 : d0; ++sp; }
                    node = first0 ? c0 : c1;
                } else if (h0) {
                    node = c0;
                } else if (h1) {
                    node = c1;
                } else {
                    pop();
                }
            }
        }
        // ---- leaves: cull, reference-order evaluation of the survivors
        if (node < 0) {
            const int32_t code = -1 - node;
            const int first = code >> 3, cnt = (code & 7) + 1;
            for (int k = 0; k < cnt; ++k) {
                const float4 ca = a.leafAB[2 * (first + k)], cb = a.leafAB[2 * (first + k) + 1];
                ++culls;
                if (cull_pass(px, py, pz, ca, cb, thr)) {
                    const float d = eval_pair<GUARD, NFMA, false>(px, py, pz, ca, cb, a.atol, a.eps, nullptr);
                    const unsigned long long kk = make_key(d, static_cast<uint32_t>(a.leaf_rows[first + k]));
                    ++pairs;
                    if (kk < key) { key = kk; thr = thr_of(key, slack); }
                }
            }
            if (static_cast<uint32_t>(key >> 32) == 0u) {       // NaN: nothing can beat it
                a.pend_keys[slot] = key;
                a.win[row] = static_cast<int32_t>(key_index(key));
                node = BVH_IDLE;
            } else {
                pop();
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

int search_bvh(tm_handle *h, const LabelArgs &a, DevStats *dst, int32_t *win) {
    BvhArgs b;
    b.pts = a.pts; b.row_stride = a.row_stride;
    b.pend_idx = h->pend_idx.as<int32_t>();
    b.pend_keys = h->keys.as<unsigned long long>();
    b.d_pending = &dst->pending;
    b.win = win;
    b.nodes = h->bvh_nodes.as<BvhNode>();
    b.leafAB = h->bvh_leafAB.as<float4>();
    b.leaf_rows = h->bvh_rows.as<int32_t>();
    b.root = h->bvh_root; b.count = h->bvh_count;
    b.recA = h->recA.as<float4>(); b.recB = h->recB.as<float4>();
    b.special = h->special.as<int32_t>(); b.aligned = h->aligned.as<int32_t>();
    b.n_special = h->n_special; b.n_aligned = h->n_aligned;
    b.atol = a.prm.perp_atol; b.eps = a.prm.norm_eps; b.maxabs = h->maxabs; b.slack_floor = h->slack_floor;
    b.st = dst;
    const bool guard = a.prm.norm_eps > 0.f, nfma = a.prm.norm_fma != 0;
    // one lane per pending point: small calls get a small grid
    const int grid = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(h->sm_count) * 12, std::max<int64_t>(h->sm_count, a.n / 2048)));
    if (guard) { if (nfma) bvh_kernel<true, true><<<grid, 128, 0, a.stream>>>(b); else bvh_kernel<true, false><<<grid, 128, 0, a.stream>>>(b); }
    else       { if (nfma) bvh_kernel<false, true><<<grid, 128, 0, a.stream>>>(b); else bvh_kernel<false, false><<<grid, 128, 0, a.stream>>>(b); }
    TM_KCHECK(h, a.stream, "bvh_kernel");
    h->stats.launches += 1;
    return TM_OK;
}

}  // namespace tmn

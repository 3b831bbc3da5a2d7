// Multi-GPU entry points of the C ABI: the point -> cylinder search shards by points, so the only
// communication is ONE broadcast of the cylinder table from the rank that read the QSM; every rank then labels its own rows
// through tm_label_points / tm_label_cloud_host.  NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch
// process that is the library torch already loaded), so the library itself has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "tm_core.cuh"

namespace tmn {

struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    char why[256] = {0};
};

static NcclApi api;

static NcclApi *nccl_api() {
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { snprintf(api.why, sizeof(api.why), "libnccl.so.2 not found: %s", dlerror()); return; }
#define TM_NCCL_SYM(field, sym)                                                                  \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, #sym));                     \
    if (!api.field) { snprintf(api.why, sizeof(api.why), "%s missing from libnccl", #sym); api.lib = nullptr; return; }
        TM_NCCL_SYM(GetUniqueId, ncclGetUniqueId)
        TM_NCCL_SYM(CommInitRank, ncclCommInitRank)
        TM_NCCL_SYM(CommInitAll, ncclCommInitAll)
        TM_NCCL_SYM(CommDestroy, ncclCommDestroy)
        TM_NCCL_SYM(Broadcast, ncclBroadcast)
        TM_NCCL_SYM(GroupStart, ncclGroupStart)
        TM_NCCL_SYM(GroupEnd, ncclGroupEnd)
        TM_NCCL_SYM(GetErrorString, ncclGetErrorString)
#undef TM_NCCL_SYM
    });
    return api.lib ? &api : nullptr;
}

static const char *nccl_why() {
    nccl_api();
    return api.why[0] ? api.why : "NCCL is not available";
}

#define TM_NCCL(h, api, expr)                                                                                      \
    do {                                                                                                           \
        ncclResult_t _r = (expr);                                                                                  \
        if (_r != ncclSuccess) return tmn::fail((h), TM_ERR_CUDA, "%s failed: %s", #expr, (api)->GetErrorString(_r)); \
    } while (0)

// (M,9) float32 rows: start xyz | axis_unit xyz | axis_length | radius | id bits — the layout sharding.pack_table uses
__global__ void pack9_kernel(const float *__restrict__ start, int64_t s_rs, int64_t s_cs, const float *__restrict__ unit, int64_t u_rs,
                             int64_t u_cs, const float *__restrict__ length, int64_t l_s, const float *__restrict__ radius, int64_t r_s,
                             const int32_t *__restrict__ ids, int64_t i_s, int64_t m, float *__restrict__ table) {
    const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (c >= m) return;
    float *row = table + 9 * c;
    row[0] = start[c * s_rs]; row[1] = start[c * s_rs + s_cs]; row[2] = start[c * s_rs + 2 * s_cs];
    row[3] = unit[c * u_rs]; row[4] = unit[c * u_rs + u_cs]; row[5] = unit[c * u_rs + 2 * u_cs];
    row[6] = length[c * l_s];
    row[7] = radius[c * r_s];
    row[8] = __int_as_float(ids ? ids[c * i_s] : static_cast<int32_t>(c));
}

static int install_table(tm_handle *h, const float *table, int64_t m, cudaStream_t st) {
    return tm_set_cylinders(h, table, 9, 1, table + 3, 9, 1, table + 6, 9, table + 7, 9, reinterpret_cast<const int32_t *>(table + 8), 9, m, st);
}

}  // namespace tmn

using namespace tmn;

extern "C" {

int tm_cylinder_count(const tm_handle *h, int64_t *m) {
    if (!h || !m) return TM_ERR_INVALID;
    *m = h->have_cyl ? h->m : 0;
    return TM_OK;
}

int tm_comm_unique_id(void *id_out) {
    if (!id_out) return TM_ERR_INVALID;
    NcclApi *api = nccl_api();
    if (!api) return TM_ERR_STATE;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return TM_ERR_CUDA;
    memcpy(id_out, &id, sizeof(id));
    return TM_OK;
}

int tm_comm_init_rank(tm_handle *h, const void *id, int32_t nranks, int32_t rank) {
    if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, TM_ERR_INVALID, "tm_comm_init_rank: bad argument%s%s");
    NcclApi *api = nccl_api();
    if (!api) return fail(h, TM_ERR_STATE, "%s%s", nccl_why());
    if (h->comm) return fail(h, TM_ERR_STATE, "tm_comm_init_rank: the handle already has a communicator%s%s");
    TM_CUDA(h, cudaSetDevice(h->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    TM_NCCL(h, api, api->CommInitRank(&comm, nranks, uid, rank));
    h->comm = comm;
    h->comm_rank = rank;
    h->comm_size = nranks;
    return TM_OK;
}

int tm_comm_init_all(tm_handle **handles, int32_t ndev) {
    if (!handles || ndev < 1 || ndev > 64) return TM_ERR_INVALID;
    for (int i = 0; i < ndev; ++i) if (!handles[i]) return TM_ERR_INVALID;
    tm_handle *h0 = handles[0];
    NcclApi *api = nccl_api();
    if (!api) return fail(h0, TM_ERR_STATE, "%s%s", nccl_why());
    int devs[64];
    ncclComm_t comms[64];
    for (int i = 0; i < ndev; ++i) {
        if (handles[i]->comm) return fail(h0, TM_ERR_STATE, "tm_comm_init_all: a handle already has a communicator%s%s");
        devs[i] = handles[i]->device;
    }
    TM_NCCL(h0, api, api->CommInitAll(comms, ndev, devs));
    for (int i = 0; i < ndev; ++i) { handles[i]->comm = comms[i]; handles[i]->comm_rank = i; handles[i]->comm_size = ndev; }
    return TM_OK;
}

int tm_comm_destroy(tm_handle *h) {
    if (!h) return TM_ERR_INVALID;
    if (!h->comm) return TM_OK;
    NcclApi *api = nccl_api();
    if (api) api->CommDestroy(static_cast<ncclComm_t>(h->comm));
    h->comm = nullptr;
    h->comm_size = 0;
    return TM_OK;
}

int tm_comm_info(const tm_handle *h, int32_t *rank, int32_t *nranks) {
    if (!h) return TM_ERR_INVALID;
    if (rank) *rank = h->comm ? h->comm_rank : 0;
    if (nranks) *nranks = h->comm ? h->comm_size : 1;
    return TM_OK;
}

// one process per GPU: every rank calls this; the cylinder arguments are read on `root` only
int tm_broadcast_cylinders(tm_handle *h, const float *start, int64_t s_rs, int64_t s_cs, const float *unit, int64_t u_rs, int64_t u_cs,
                           const float *length, int64_t l_s, const float *radius, int64_t r_s, const int32_t *ids, int64_t i_s,
                           int64_t m, int32_t root, void *stream) {
    if (!h) return TM_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TM_CUDA(h, cudaSetDevice(h->device));
    const bool single = !h->comm || h->comm_size == 1;
    const bool is_root = single || h->comm_rank == root;
    if (!single && (root < 0 || root >= h->comm_size)) return fail(h, TM_ERR_INVALID, "tm_broadcast_cylinders: bad root%s%s");
    if (is_root && (m < 0 || (m > 0 && (!start || !unit || !length || !radius))))
        return fail(h, TM_ERR_INVALID, "tm_broadcast_cylinders: bad argument on the root%s%s");
    NcclApi *api = single ? nullptr : nccl_api();
    if (!single && !api) return fail(h, TM_ERR_STATE, "%s%s", nccl_why());
    // the row count travels first (8 bytes), then the packed table
    TM_CUDA(h, h->comm_count.ensure(sizeof(long long)));
    long long count = is_root ? static_cast<long long>(m) : 0;
    if (!single) {
        TM_CUDA(h, cudaMemcpyAsync(h->comm_count.p, &count, sizeof(count), cudaMemcpyHostToDevice, st));
        TM_NCCL(h, api, api->Broadcast(h->comm_count.p, h->comm_count.p, 1, ncclInt64, root, static_cast<ncclComm_t>(h->comm), st));
        TM_CUDA(h, cudaMemcpyAsync(&count, h->comm_count.p, sizeof(count), cudaMemcpyDeviceToHost, st));
        TM_CUDA(h, cudaStreamSynchronize(st));
    }
    if (count < 0 || count > 0x7fffffffLL) return fail(h, TM_ERR_INVALID, "tm_broadcast_cylinders: bad row count from the root%s%s");
    TM_CUDA(h, h->comm_table.ensure(sizeof(float) * 9 * static_cast<size_t>(std::max<long long>(count, 1))));
    float *table = h->comm_table.as<float>();
    if (is_root && count > 0) {
        pack9_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, st>>>(start, s_rs, s_cs, unit, u_rs, u_cs, length, l_s, radius, r_s,
                                                                                 ids, i_s, count, table);
        TM_KCHECK(h, st, "pack9_kernel");
    }
    if (!single && count > 0)
        TM_NCCL(h, api, api->Broadcast(table, table, static_cast<size_t>(count) * 9, ncclFloat32, root, static_cast<ncclComm_t>(h->comm), st));
    return install_table(h, table, count, st);
}

// one process, several devices (tm_comm_init_all): handles[root_index] holds the table's source pointers
int tm_broadcast_cylinders_all(tm_handle **handles, int32_t ndev, const float *start, int64_t s_rs, int64_t s_cs, const float *unit,
                               int64_t u_rs, int64_t u_cs, const float *length, int64_t l_s, const float *radius, int64_t r_s,
                               const int32_t *ids, int64_t i_s, int64_t m, int32_t root_index) {
    if (!handles || ndev < 1 || root_index < 0 || root_index >= ndev || m < 0) return TM_ERR_INVALID;
    tm_handle *hr = handles[root_index];
    if (!hr) return TM_ERR_INVALID;
    if (ndev == 1) return tm_broadcast_cylinders(hr, start, s_rs, s_cs, unit, u_rs, u_cs, length, l_s, radius, r_s, ids, i_s, m, 0, nullptr);
    NcclApi *api = nccl_api();
    if (!api) return fail(hr, TM_ERR_STATE, "%s%s", nccl_why());
    for (int i = 0; i < ndev; ++i) {
        tm_handle *h = handles[i];
        if (!h || !h->comm || h->comm_size != ndev) return fail(hr, TM_ERR_STATE, "tm_broadcast_cylinders_all: call tm_comm_init_all first%s%s");
        TM_CUDA(h, cudaSetDevice(h->device));
        TM_CUDA(h, h->comm_table.ensure(sizeof(float) * 9 * static_cast<size_t>(std::max<int64_t>(m, 1))));
    }
    TM_CUDA(hr, cudaSetDevice(hr->device));
    if (m > 0) {
        pack9_kernel<<<static_cast<unsigned>((m + 255) / 256), 256>>>(start, s_rs, s_cs, unit, u_rs, u_cs, length, l_s, radius, r_s, ids, i_s, m,
                                                                      hr->comm_table.as<float>());
        TM_CUDA(hr, cudaGetLastError());
        TM_CUDA(hr, cudaDeviceSynchronize());
        TM_NCCL(hr, api, api->GroupStart());
        for (int i = 0; i < ndev; ++i) {
            tm_handle *h = handles[i];
            cudaSetDevice(h->device);
            ncclResult_t r = api->Broadcast(h->comm_table.p, h->comm_table.p, static_cast<size_t>(m) * 9, ncclFloat32, root_index,
                                            static_cast<ncclComm_t>(h->comm), nullptr);
            if (r != ncclSuccess) { api->GroupEnd(); return fail(hr, TM_ERR_CUDA, "ncclBroadcast failed: %s%s", api->GetErrorString(r)); }
        }
        TM_NCCL(hr, api, api->GroupEnd());
    }
    for (int i = 0; i < ndev; ++i) {
        tm_handle *h = handles[i];
        cudaSetDevice(h->device);
        const int rc = install_table(h, h->comm_table.as<float>(), m, nullptr);
        if (rc != TM_OK) return rc;
    }
    return TM_OK;
}

}  // extern "C"

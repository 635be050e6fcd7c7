// Multi-GPU: one process per GPU, atoms partitioned by ownership ranges of the spatially sorted
// order (whole molecules, Morton-contiguous => spatially compact domains), NCCL over NVLink.
//
// The reference is single-process; this is the engine's own decomposition:
//   * every rank keeps a full copy of the positions; before each pair-force evaluation the owned
//     segments are exchanged (grouped ncclBroadcast = all-gather with uneven segments).  With the
//     full (both-directions) neighbour list there is NO force return;
//   * velocities, forces, thermostat variables, neighbour lists and all integration work exist
//     only for the owned range; bonded terms are intramolecular and need no communication;
//   * global sums (mvv, energies) are ncclAllReduce'd; the skin test runs on the replicated
//     positions, so every rank takes the same rebuild decision without communication;
//   * all NCCL calls are enqueued on the context stream and are captured into the per-step graph.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "ctx.h"

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl(b2_context* ctx) {
    if (g_nccl.handle) return B2_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "cannot load libnccl.so.2: %s", dlerror());
#define LOAD(field, symbol)                                                                  \
    *(void**)(&g_nccl.field) = dlsym(h, symbol);                                             \
    if (!g_nccl.field) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "libnccl.so.2 lacks %s", symbol);
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(Broadcast, "ncclBroadcast")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    g_nccl.handle = h;
    return B2_OK;
}

#define B2_NCCL(call)                                                                          \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != ncclSuccess)                                                                 \
            return b2_fail(ctx, B2_ERR_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

extern "C" int b2_comm_unique_id(char* out128) {
    B2_TRY(load_nccl(nullptr));
    ncclUniqueId id;
    b2_context* ctx = nullptr;
    B2_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(out128, &id, 128);
    return B2_OK;
}

extern "C" int b2_comm_init(b2_context* ctx, int nranks, int rank, const char* id128) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return b2_fail(ctx, B2_ERR_ARG, "bad communicator arguments");
    if (ctx->pme_forces.size() > 0) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "PME is not available with domain decomposition yet");
    if (nranks == 1) { ctx->nranks = 1; ctx->rank = 0; return B2_OK; }
    B2_TRY(load_nccl(ctx));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    B2_CUDA(cudaSetDevice(ctx->device));
    ncclComm_t comm;
    B2_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    ctx->have_order = false;
    program_release(ctx);
    return B2_OK;
}

void dist_release(b2_context* ctx) {
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
}

// ownership ranges: split the sorted order at molecule boundaries nearest to k*n/P
extern "C" int b2_partition_ranges(int n, const int* molecule_sorted, int nranks, int* out_ranges) {
    if (n < 0 || nranks < 1 || !out_ranges || (n > 0 && !molecule_sorted)) return B2_ERR_ARG;
    const int P = nranks;
    out_ranges[0] = 0;
    out_ranges[P] = n;
    for (int k = 1; k < P; k++) {
        const int target = (int)((long long)n*k/P);
        int up = target, down = target;
        while (up < n && up > 0 && molecule_sorted[up] == molecule_sorted[up-1]) up++;
        while (down > 0 && down < n && molecule_sorted[down] == molecule_sorted[down-1]) down--;
        int cut = (up - target <= target - down) ? up : down;
        if (cut < out_ranges[k-1]) cut = out_ranges[k-1];
        out_ranges[k] = cut;
    }
    return B2_OK;
}

int dist_partition(b2_context* ctx) {
    const int n = ctx->n, P = ctx->nranks;
    std::vector<int> mol(n);
    for (int s = 0; s < n; s++) mol[s] = ctx->h_mol[ctx->h_orig[s]];
    ctx->range.assign(P + 1, 0);
    b2_partition_ranges(n, mol.data(), P, ctx->range.data());
    ctx->a_lo = ctx->range[ctx->rank];
    ctx->a_hi = ctx->range[ctx->rank + 1];
    ctx->g_lo = ctx->a_lo/B2_GROUP;
    ctx->g_hi = (ctx->a_hi + B2_GROUP - 1)/B2_GROUP;
    if (ctx->a_hi == ctx->a_lo) { ctx->g_lo = ctx->g_hi = 0; }
    return B2_OK;
}

extern "C" int b2_comm_info(b2_context* ctx, int* rank, int* nranks, int* lo, int* hi, long long* exchanges) {
    if (!ctx) return B2_ERR_ARG;
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    if (lo) *lo = ctx->a_lo;
    if (hi) *hi = ctx->a_hi;
    if (exchanges) *exchanges = ctx->counters[7];
    return B2_OK;
}

// all ranks end up with every rank's owned segment of a [n][width] array of `bytes`-byte elements
static int allgather_segments(b2_context* ctx, void* base, size_t elem_bytes) {
    if (ctx->nranks == 1) return B2_OK;
    B2_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < ctx->nranks; r++) {
        const size_t lo = ctx->range[r], count = ctx->range[r+1] - ctx->range[r];
        if (count == 0) continue;
        char* p = (char*)base + lo*elem_bytes;
        B2_NCCL(g_nccl.Broadcast(p, p, count*elem_bytes, ncclChar, r, (ncclComm_t)ctx->comm, ctx->stream));
    }
    B2_NCCL(g_nccl.GroupEnd());
    ctx->counters[7]++;
    return B2_OK;
}

int dist_sync_positions(b2_context* ctx) {
    if (ctx->nranks == 1 || ctx->x_synced == ctx->pos_version) return B2_OK;
    B2_TRY(allgather_segments(ctx, ctx->x, 3*sizeof(double)));
    ctx->x_synced = ctx->pos_version;
    return B2_OK;
}

int dist_gather3(b2_context* ctx, double* array) { return allgather_segments(ctx, array, 3*sizeof(double)); }
int dist_gather_forces(b2_context* ctx, float4* array) { return allgather_segments(ctx, array, sizeof(float4)); }

int dist_allreduce(b2_context* ctx, double* values, int count) {
    if (ctx->nranks == 1) return B2_OK;
    B2_NCCL(g_nccl.AllReduce(values, values, count, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return B2_OK;
}

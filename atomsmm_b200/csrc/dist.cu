// Multi-GPU: one process per GPU, atoms partitioned by ownership ranges of the spatially sorted
// order (whole molecules, Morton-contiguous => spatially compact domains), NCCL over NVLink.
//
// The reference is single-process; this is the engine's own decomposition:
//   * velocities, forces, thermostat variables, neighbour lists and all integration work exist
//     only for the owned range; bonded terms are intramolecular and need no communication; with the
//     full (both-directions) neighbour list there is NO force return;
//   * HALO EXCHANGE over peer memory (default): every rank maps its peers' position arrays
//     (cudaIpc, NVLink).  Before a pair-force evaluation each rank tests the skin criterion on its
//     OWN atoms, posts "positions final + rebuild wanted?" to all peers, and pulls the positions of
//     its halo -- the j-groups its last list build saw and does not own, ~5 MB instead of the
//     100 MB of a full replica at 4.2 M atoms -- straight out of the owners' memory; if any rank
//     wants a rebuild, everybody pulls everything and rebuilds (same decision everywhere, taken on
//     the device, so the whole step stays one CUDA graph).  Sums are pushed into the peers'
//     signal blocks and added in rank order (dd.cuh);
//   * fall-back (B2_DD_EXCHANGE=nccl or no peer access): every rank keeps a full replica, owned
//     segments are all-gathered with grouped ncclBroadcast before each pair-force evaluation and
//     sums are ncclAllReduce'd;
//   * host-initiated gathers of the complete state (getters, re-ordering) always use NCCL.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "ctx.h"
#include "dd.cuh"

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
    if (ctx->p2p) {
        for (int r = 0; r < ctx->nranks; r++) {
            if (r == ctx->rank) continue;
            if (ctx->peer_x[r]) cudaIpcCloseMemHandle(ctx->peer_x[r]);
            if (ctx->peer_sig[r]) cudaIpcCloseMemHandle(ctx->peer_sig[r]);
        }
        ctx->p2p = false;
    }
    cudaFree(ctx->sig); cudaFree(ctx->dd_state); cudaFree(ctx->halo_mark); cudaFree(ctx->halo_groups); cudaFree(ctx->halo_count);
    ctx->sig = nullptr; ctx->dd_state = nullptr; ctx->halo_mark = nullptr; ctx->halo_groups = nullptr; ctx->halo_count = nullptr;
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

// every rank ends up with the complete, current position array (NCCL all-gather of the owned segments)
int dist_sync_positions(b2_context* ctx) {
    if (ctx->nranks == 1 || ctx->x_synced == ctx->pos_version) return B2_OK;
    B2_TRY(dist_before_move(ctx));        // x is about to be overwritten outside the owned range: no reader may be active
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

// ---------------------------------------------------------------------------------------------
// peer-memory exchange (see dd.cuh)
// ---------------------------------------------------------------------------------------------
static DDPeers make_peers(const b2_context* ctx) {
    DDPeers P;
    memset(&P, 0, sizeof(P));
    P.rank = ctx->rank; P.nranks = ctx->nranks;
    for (int r = 0; r <= ctx->nranks && r <= B2_MAX_RANKS; r++) P.range[r] = r < (int)ctx->range.size() ? ctx->range[r] : ctx->n;
    for (int r = 0; r < ctx->nranks; r++) { P.x[r] = ctx->peer_x[r]; P.sig[r] = ctx->peer_sig[r]; }
    return P;
}

void dd_fill_peers(const b2_context* ctx, void* out) { *(DDPeers*)out = make_peers(ctx); }

// "my owned positions of this exchange are final", plus this rank's verdict of the skin test
__global__ void k_dd_post(DDPeers P, unsigned long long* state, const int* __restrict__ nl_flags) {
    __shared__ unsigned long long epoch;
    if (threadIdx.x == 0) { epoch = state[0] + 1ull; state[0] = epoch; }
    __syncthreads();
    const int r = threadIdx.x;
    if (r < P.nranks && r != P.rank) {
        const unsigned long long flag = nl_flags[0] != 0 ? 1ull : 0ull;
        __threadfence_system();
        dd_st_release(&P.sig[r][DD_POST + P.rank], (epoch << 1) | flag);
    }
}

#define DD_TILE 21            // groups per block and pass: 21 x 12 words = 252 of 256 threads busy
// wait for every peer's post, then copy the halo (or, when anybody wants a rebuild, all foreign
// atoms) out of the owners' memory; the last block to finish publishes the common rebuild
// decision, resets the halo counter for the coming rebuild and acknowledges to the owners
__global__ void __launch_bounds__(256) k_dd_pull(DDPeers P, unsigned long long* state, int* nl_flags, double* __restrict__ x,
                                                 int4* __restrict__ xq, double sx, double sy, double sz,
                                                 int n, const int* __restrict__ halo_groups, int* halo_count) {
    __shared__ int s_flag;
    __shared__ unsigned long long s_epoch;
    __shared__ bool last;
    if (threadIdx.x == 0) { s_flag = nl_flags[0] != 0 ? 1 : 0; s_epoch = state[0]; }
    const unsigned long long t_start = dd_now();
    __syncthreads();
    const unsigned long long e = s_epoch;
    if ((int)threadIdx.x < P.nranks && (int)threadIdx.x != P.rank) {
        const unsigned long long v = dd_wait(&P.sig[P.rank][DD_POST + threadIdx.x], 1, e, state);
        if (v & 1ull) atomicOr(&s_flag, 1);
    }
    __syncthreads();
    // exchange clock (b2_comm_timing): block 0 saw the whole wait; the last block closes the copy
    if (blockIdx.x == 0 && threadIdx.x == 0) { state[4] += dd_now() - t_start; state[7] = dd_now(); }
    // COALESCED copy.  A group of 8 atoms is 192 contiguous, 64-byte aligned bytes = twelve 16-byte words; a block
    // takes DD_TILE groups per pass and thread t copies word t % 12 of group t / 12, so that consecutive lanes read
    // consecutive addresses of the owner's memory (round 2c gave every thread a 48-byte atom pair: three loads with
    // a 48-byte stride between lanes, i.e. every 128-byte line crossed NVLink three times -- 124 GB/s of useful
    // data).  Two passes are in flight per thread.  A word can straddle two atoms with different owners at an
    // ownership boundary (cuts fall on molecules, not on groups): then its two doubles are handled one by one.
    // After a barrier the block converts the atoms it has just copied to the fixed-point copy the pair tiles read.
    // Halo mode walks the pull list, rebuild mode all groups.
    const int ngroups_all = (n + B2_GROUP - 1)/B2_GROUP;
    const int units = s_flag ? ngroups_all : *halo_count;
    const long long ndoubles = 3ll*n;
    const int t = threadIdx.x, gi = t/12, w = t - 12*gi;
    for (int base = blockIdx.x*DD_TILE*2; base < units; base += gridDim.x*DD_TILE*2) {
        double2 val[2];
        long long d0[2];
        int own[2][2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int unit = base + u*DD_TILE + gi;
            d0[u] = -1;
            own[u][0] = own[u][1] = P.rank;
            if (gi < DD_TILE && unit < units) {
                const int g = s_flag ? unit : halo_groups[unit];
                const long long d = 24ll*g + 2*w;
                if (d < ndoubles) {
                    d0[u] = d;
                    own[u][0] = dd_owner(P, (int)(d/3));
                    own[u][1] = d + 1 < ndoubles ? dd_owner(P, (int)((d + 1)/3)) : P.rank;
                    if (own[u][0] == own[u][1]) {
                        if (own[u][0] != P.rank) val[u] = __ldcg(reinterpret_cast<const double2*>(P.x[own[u][0]] + d));
                    } else {
                        if (own[u][0] != P.rank) val[u].x = __ldcg(P.x[own[u][0]] + d);
                        if (own[u][1] != P.rank) val[u].y = __ldcg(P.x[own[u][1]] + d + 1);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (d0[u] < 0) continue;
            if (own[u][0] == own[u][1]) {
                if (own[u][0] != P.rank) *reinterpret_cast<double2*>(x + d0[u]) = val[u];
            } else {
                if (own[u][0] != P.rank) x[d0[u]] = val[u].x;
                if (own[u][1] != P.rank) x[d0[u] + 1] = val[u].y;
            }
        }
        __syncthreads();
        // fixed-point copy of the atoms of this pass: 2 x DD_TILE groups x 8 atoms
        for (int k = t; k < 2*DD_TILE*B2_GROUP; k += blockDim.x) {
            const int unit = base + k/B2_GROUP;
            if (unit >= units) continue;
            const int g = s_flag ? unit : halo_groups[unit];
            const int a = g*B2_GROUP + (k & (B2_GROUP - 1));
            if (a >= n || dd_owner(P, a) == P.rank) continue;
            xq[a] = make_int4(b2_to_fixed(x[3ll*a], sx), b2_to_fixed(x[3ll*a+1], sy), b2_to_fixed(x[3ll*a+2], sz), 0);
        }
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd((unsigned*)(state + 3), 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) {
        *(unsigned*)(state + 3) = 0u;
        nl_flags[0] = s_flag;
        if (s_flag) *halo_count = 0;
        const unsigned long long t_copy = *(volatile unsigned long long*)(state + 7);
        const unsigned long long t_end = dd_now();
        if (t_copy != 0ull && t_end > t_copy) state[5] += t_end - t_copy;
        state[6] += 1ull;
    }
    if ((int)threadIdx.x < P.nranks && (int)threadIdx.x != P.rank) {
        __threadfence_system();
        dd_st_release(&P.sig[threadIdx.x][DD_ACK + P.rank], e);
    }
}

// an owner may move its atoms again only after every peer has finished reading them
__global__ void k_dd_wait_acks(DDPeers P, unsigned long long* state, unsigned long long* ack_ns) {
    const unsigned long long e = state[0];
    const unsigned long long t0 = dd_now();
    if ((int)threadIdx.x < P.nranks && (int)threadIdx.x != P.rank) dd_wait(&P.sig[P.rank][DD_ACK + threadIdx.x], 0, e, state);
    __syncwarp();
    if (threadIdx.x == 0) *ack_ns += dd_now() - t0;
}

__global__ void k_halo_compact(int ngroups, unsigned char* __restrict__ mark, int* __restrict__ halo_groups, int* halo_count,
                               const int* __restrict__ flags) {
    if (!flags[0]) return;
    const int g = blockIdx.x*blockDim.x + threadIdx.x;
    if (g >= ngroups || !mark[g]) return;
    mark[g] = 0;
    halo_groups[atomicAdd(halo_count, 1)] = g;
}

// globals[target] <- sum over ranks of globals[target]
__global__ void k_dd_reduce(DDPeers P, unsigned long long* state, double* value) {
    __shared__ double part[B2_MAX_RANKS];
    __shared__ unsigned long long epoch;
    const double total = dd_allreduce_sum(P, state, *value, threadIdx.x, part, &epoch);
    __syncthreads();
    if (threadIdx.x == 0) *value = total;
}

int dist_exchange_halo(b2_context* ctx) {
    if (!ctx->p2p) return B2_OK;
    DDPeers P = make_peers(ctx);
    k_dd_post<<<1, 32, 0, ctx->stream>>>(P, ctx->dd_state, ctx->nl_flags);
    B2_LAUNCH_CHECK();
    k_dd_pull<<<592, 256, 0, ctx->stream>>>(P, ctx->dd_state, ctx->nl_flags, ctx->x, ctx->xq, 4294967296.0/ctx->box[0],
                                            4294967296.0/ctx->box[1], 4294967296.0/ctx->box[2], ctx->n, ctx->halo_groups,
                                            ctx->halo_count);
    B2_LAUNCH_CHECK();
    ctx->acks_pending = true;
    ctx->counters[7]++;
    return B2_OK;
}

int dist_before_move(b2_context* ctx) {
    if (!ctx->p2p || !ctx->acks_pending) return B2_OK;
    DDPeers P = make_peers(ctx);
    k_dd_wait_acks<<<1, 32, 0, ctx->stream>>>(P, ctx->dd_state, ctx->dd_state + 8);
    B2_LAUNCH_CHECK();
    ctx->acks_pending = false;
    return B2_OK;
}

int dist_halo_compact(b2_context* ctx) {
    if (!ctx->p2p) return B2_OK;
    const int T = 256;
    k_halo_compact<<<(ctx->ngroups + T - 1)/T, T, 0, ctx->stream>>>(ctx->ngroups, ctx->halo_mark, ctx->halo_groups,
                                                                    ctx->halo_count, ctx->nl_flags);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

// sum of one device double over the ranks, on the context stream
int dist_reduce_value(b2_context* ctx, double* value) {
    if (ctx->nranks == 1) return B2_OK;
    if (!ctx->p2p) return dist_allreduce(ctx, value, 1);
    DDPeers P = make_peers(ctx);
    k_dd_reduce<<<1, 32, 0, ctx->stream>>>(P, ctx->dd_state, value);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

// ---- set-up: exchange of cudaIpc handles (host side does the all-gather, e.g. torch.distributed) ---
struct DDHandles {
    cudaIpcMemHandle_t x, sig;
    int device, n;
};
static_assert(sizeof(DDHandles) <= 256, "handle record must fit the ABI's 256 bytes");

extern "C" int b2_comm_export(b2_context* ctx, char* out256) {
    if (!ctx || !out256 || ctx->n == 0) return b2_fail(ctx, B2_ERR_STATE, "set particles before exporting the peer handles");
    B2_CUDA(cudaSetDevice(ctx->device));
    if (ctx->sig == nullptr) {
        // 2 MB: an allocation of its own (small cudaMalloc blocks share a slab, and a slab can be opened only
        // once per peer -- the position array of a small system may live in one)
        B2_CUDA(cudaMalloc(&ctx->sig, 2u << 20));
        B2_CUDA(cudaMemset(ctx->sig, 0, 2u << 20));
    }
    DDHandles h;
    memset(&h, 0, sizeof(h));
    B2_CUDA(cudaIpcGetMemHandle(&h.x, ctx->x));
    B2_CUDA(cudaIpcGetMemHandle(&h.sig, ctx->sig));
    h.device = ctx->device; h.n = ctx->n;
    memset(out256, 0, 256);
    memcpy(out256, &h, sizeof(h));
    return B2_OK;
}

// all256: nranks records as written by b2_comm_export, in rank order.  On failure (no peer access, IPC not
// permitted) the context stays in the NCCL all-gather mode and the reason is in b2_last_error.
extern "C" int b2_comm_import(b2_context* ctx, int nranks, const char* all256) {
    if (ctx && nranks == -1) {                       // "not every rank could map its peers": back to the NCCL mode
        ctx->p2p = false;
        program_release(ctx);
        return B2_OK;
    }
    if (!ctx || !all256 || nranks != ctx->nranks) return b2_fail(ctx, B2_ERR_ARG, "bad peer handle table");
    if (nranks == 1) return B2_OK;
    if (nranks > B2_MAX_RANKS) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "at most %d ranks in peer-memory mode", B2_MAX_RANKS);
    B2_CUDA(cudaSetDevice(ctx->device));
    for (int r = 0; r < nranks; r++) {
        DDHandles h;
        memcpy(&h, all256 + 256*(size_t)r, sizeof(h));
        if (h.n != ctx->n) return b2_fail(ctx, B2_ERR_ARG, "rank %d describes %d atoms, this rank %d", r, h.n, ctx->n);
        if (r == ctx->rank) { ctx->peer_x[r] = ctx->x; ctx->peer_sig[r] = ctx->sig; continue; }
        int can = 0;
        cudaDeviceCanAccessPeer(&can, ctx->device, h.device);
        if (!can) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "device %d cannot access device %d", ctx->device, h.device);
        void *px = nullptr, *ps = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&px, h.x, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&ps, h.sig, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return b2_fail(ctx, B2_ERR_UNSUPPORTED, "cudaIpcOpenMemHandle for rank %d failed: %s", r, cudaGetErrorString(e));
        }
        ctx->peer_x[r] = (double*)px;
        ctx->peer_sig[r] = (unsigned long long*)ps;
    }
    if (ctx->dd_state == nullptr) {
        B2_CUDA(cudaMalloc(&ctx->dd_state, sizeof(unsigned long long)*16));
        B2_CUDA(cudaMemset(ctx->dd_state, 0, sizeof(unsigned long long)*16));
        B2_CUDA(cudaMalloc(&ctx->halo_mark, ctx->ngroups));
        B2_CUDA(cudaMemset(ctx->halo_mark, 0, ctx->ngroups));
        B2_CUDA(cudaMalloc(&ctx->halo_groups, sizeof(int)*ctx->ngroups));
        B2_CUDA(cudaMalloc(&ctx->halo_count, sizeof(int)));
        B2_CUDA(cudaMemset(ctx->halo_count, 0, sizeof(int)));
    }
    ctx->p2p = true;
    program_release(ctx);
    return B2_OK;
}

extern "C" int b2_comm_timing(b2_context* ctx, double out[4]) {
    if (!ctx || !out) return B2_ERR_ARG;
    out[0] = out[1] = out[2] = out[3] = 0.0;
    if (!ctx->p2p || !ctx->dd_state) return B2_OK;
    unsigned long long h[16];
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    B2_CUDA(cudaMemcpy(h, ctx->dd_state, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] = 1e-9*(double)h[4]; out[1] = 1e-9*(double)h[5]; out[2] = (double)h[6]; out[3] = 1e-9*(double)h[8];
    return B2_OK;
}

extern "C" int b2_comm_mode(b2_context* ctx, int* peer_memory, long long* halo_atoms) {
    if (!ctx) return B2_ERR_ARG;
    if (peer_memory) *peer_memory = ctx->p2p ? 1 : 0;
    if (halo_atoms) {
        *halo_atoms = 0;
        if (ctx->p2p && ctx->halo_count) {
            int c = 0;
            B2_CUDA(cudaStreamSynchronize(ctx->stream));
            B2_CUDA(cudaMemcpy(&c, ctx->halo_count, sizeof(int), cudaMemcpyDeviceToHost));
            *halo_atoms = (long long)c*B2_GROUP;
        }
    }
    return B2_OK;
}

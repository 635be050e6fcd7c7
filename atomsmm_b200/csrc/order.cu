// Device-side spatial ordering (kernel K1 of SURVEY 2.1: sort + re-sort).
//
// The engine keeps whole molecules contiguous and sorts them along a Hilbert curve (api.cu).  Round 1 computed the
// order on the host: download of all positions, one Hilbert key per molecule, std::sort, O(N) host loops that
// permute every static per-atom table, re-upload -- 0.5 s at 4.2 M atoms, paid every few picoseconds of a long run.
// Here the whole chain runs on the device:
//     k_mol_keys          Hilbert key of every molecule (position of its first atom, caller order)
//     cub::DeviceRadixSort stable sort of (key, molecule id)      -- plain library sort, not a hot-path kernel
//     cub::DeviceScan     offsets of the sorted molecules
//     k_fill_order        orig / inv / molecule table in the engine's order
//     k_gather_static     masses, parameter sets (fp32 tile form + fp64), exclusion masks gathered through orig
//     k_excl_span         largest index distance between the two atoms of any exclusion
// Only `orig` (4 B per atom) travels back to the host, for the tables that are still built there (molecule chunks of
// the fused inner loop, constraint clusters, ownership ranges).  The keys are bit-identical to b2_hilbert_index (the
// host function the CPU tests pin): the same integer arithmetic, and the one floating-point expression is written
// with explicitly rounded operations so that the device compiler cannot contract it.
#include <cub/cub.cuh>

#include "ctx.h"

__host__ __device__ static inline unsigned long long hilbert_key_hd(const unsigned cell[3], int bits) {
    unsigned X[3] = {cell[0], cell[1], cell[2]};
    const unsigned M = 1u << (bits - 1);
    for (unsigned Q = M; Q > 1; Q >>= 1) {
        const unsigned P = Q - 1;
        for (int i = 0; i < 3; i++) {
            if (X[i] & Q) X[0] ^= P;
            else { const unsigned t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    for (int i = 1; i < 3; i++) X[i] ^= X[i-1];
    unsigned t = 0;
    for (unsigned Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1;
    for (int i = 0; i < 3; i++) X[i] ^= t;
    unsigned long long key = 0;
    for (int q = bits - 1; q >= 0; q--)
        for (int i = 0; i < 3; i++) key = (key << 1) | ((X[i] >> q) & 1u);
    return key;
}

int order_hilbert_bits(const double box[3]) {
    const double longest = std::max(box[0], std::max(box[1], box[2]));
    int bits = 1;
    while (bits < 20 && longest/(double)(1u << bits) > 0.32) bits++;
    return bits;
}

__global__ void k_mol_keys(int nmol, const int* __restrict__ mol_ptr, const int* __restrict__ mol_atoms,
                           const double* __restrict__ x_user, double bx, double by, double bz, int bits,
                           unsigned long long* __restrict__ keys, int* __restrict__ ids) {
    const int m = blockIdx.x*blockDim.x + threadIdx.x;
    if (m >= nmol) return;
    ids[m] = m;
    if (mol_ptr[m+1] == mol_ptr[m]) { keys[m] = ~0ull; return; }       // unused molecule id: sorts last, owns no atom
    const int a = mol_atoms[mol_ptr[m]];
    const double box[3] = {bx, by, bz};
    unsigned c[3];
    for (int d = 0; d < 3; d++) {
        const double L = box[d], p = x_user[3*(size_t)a + d];
        // w = p - L*floor(p/L), k = (int)(w/L * 2^bits): the operation order of b2_hilbert_index, no contraction
        const double w = __dadd_rn(p, -__dmul_rn(L, floor(__ddiv_rn(p, L))));
        const int k = (int)__dmul_rn(__ddiv_rn(w, L), (double)(1u << bits));
        c[d] = (unsigned)min(max(k, 0), (1 << bits) - 1);
    }
    keys[m] = hilbert_key_hd(c, bits);
}

__global__ void k_mol_sizes(int nmol, const int* __restrict__ sorted_ids, const int* __restrict__ mol_ptr, int* __restrict__ sizes) {
    const int p = blockIdx.x*blockDim.x + threadIdx.x;
    if (p >= nmol) return;
    const int m = sorted_ids[p];
    sizes[p] = mol_ptr[m+1] - mol_ptr[m];
}

// one thread per molecule of the sorted sequence (molecules are small; a large one costs one thread a longer loop)
__global__ void k_fill_order(int nmol, int n, int nmol_used, const int* __restrict__ sorted_ids, const int* __restrict__ offsets,
                             const int* __restrict__ mol_ptr, const int* __restrict__ mol_atoms, int* __restrict__ orig,
                             int* __restrict__ inv, int* __restrict__ mol_start) {
    const int p = blockIdx.x*blockDim.x + threadIdx.x;
    if (p >= nmol) return;
    const int m = sorted_ids[p], first = offsets[p];
    const int lo = mol_ptr[m], size = mol_ptr[m+1] - lo;
    for (int k = 0; k < size; k++) {
        const int a = mol_atoms[lo + k];
        orig[first + k] = a;
        inv[a] = first + k;
    }
    if (p < nmol_used) mol_start[p] = first;
    if (p == nmol_used - 1) mol_start[nmol_used] = n;
}

struct GatherArgs {
    int nsets;
    const double* sets_user[B2_MAX_SETS];
    float4* par[B2_MAX_SETS];
    double* pard[B2_MAX_SETS];
};

__global__ void k_gather_static(int n, const int* __restrict__ orig, const double* __restrict__ mass_user,
                                const unsigned long long* __restrict__ exmask_user, double* __restrict__ massd,
                                float* __restrict__ invm, unsigned long long* __restrict__ exmask, GatherArgs g) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int o = orig[s];
    const double m = mass_user[o];
    massd[s] = m;
    invm[s] = m > 0 ? (float)(1.0/m) : 0.f;
    exmask[s] = exmask_user[o];
    for (int k = 0; k < g.nsets; k++) {
        const double* src = g.sets_user[k] + 3*(size_t)o;
        const double q = src[0], sig = src[1], eps = src[2];
        g.par[k][s] = make_float4((float)q, (float)(0.5*sig), (float)sqrt(eps), 0.f);
        g.pard[k][3*(size_t)s] = q; g.pard[k][3*(size_t)s+1] = sig; g.pard[k][3*(size_t)s+2] = eps;
    }
}

__global__ void k_excl_span(int npairs, const int* __restrict__ pairs, const int* __restrict__ inv, int* __restrict__ span) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= npairs) return;
    atomicMax(span, abs(inv[pairs[2*k]] - inv[pairs[2*k+1]]));
}

// caller-order tables the device ordering gathers from; refreshed when the description changes
static int order_upload_tables(b2_context* ctx) {
    const int n = ctx->n;
    OrderDevice& D = ctx->order;
    if (ctx->h_mol_ptr.empty()) {
        // molecules in caller order: a static CSR table, built once
        int nmol = 0;
        for (int i = 0; i < n; i++) {
            if (ctx->h_mol[i] < 0) return b2_fail(ctx, B2_ERR_ARG, "negative molecule id");
            nmol = std::max(nmol, ctx->h_mol[i] + 1);
        }
        ctx->h_mol_ptr.assign(nmol + 1, 0);
        for (int i = 0; i < n; i++) ctx->h_mol_ptr[ctx->h_mol[i] + 1]++;
        for (int m = 0; m < nmol; m++) ctx->h_mol_ptr[m+1] += ctx->h_mol_ptr[m];
        ctx->h_mol_atoms.resize(n);
        std::vector<int> cursor(ctx->h_mol_ptr.begin(), ctx->h_mol_ptr.end() - 1);
        for (int i = 0; i < n; i++) ctx->h_mol_atoms[cursor[ctx->h_mol[i]]++] = i;
        D.valid = false;
    }
    if (D.valid && D.nsets == (int)ctx->h_sets.size() && D.nexcl == (int)ctx->h_excl.size()/2) return B2_OK;
    const int nmol = (int)ctx->h_mol_ptr.size() - 1;
    order_release(ctx);
    D.nmol = nmol;
    D.nmol_used = 0;
    for (int m = 0; m < nmol; m++) D.nmol_used += ctx->h_mol_ptr[m+1] > ctx->h_mol_ptr[m] ? 1 : 0;
    B2_CUDA(cudaMalloc(&D.mol_ptr, sizeof(int)*(nmol + 1)));
    B2_CUDA(cudaMalloc(&D.mol_atoms, sizeof(int)*std::max(1, n)));
    B2_CUDA(cudaMalloc(&D.keys, sizeof(unsigned long long)*2*std::max(1, nmol)));
    B2_CUDA(cudaMalloc(&D.ids, sizeof(int)*2*std::max(1, nmol)));
    B2_CUDA(cudaMalloc(&D.sizes, sizeof(int)*(nmol + 1)));
    B2_CUDA(cudaMalloc(&D.offsets, sizeof(int)*(nmol + 1)));
    B2_CUDA(cudaMalloc(&D.mass_user, sizeof(double)*n));
    B2_CUDA(cudaMalloc(&D.exmask_user, sizeof(unsigned long long)*n));
    B2_CUDA(cudaMalloc(&D.span, sizeof(int)));
    B2_CUDA(cudaMemcpy(D.mol_ptr, ctx->h_mol_ptr.data(), sizeof(int)*(nmol + 1), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(D.mol_atoms, ctx->h_mol_atoms.data(), sizeof(int)*n, cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(D.mass_user, ctx->h_mass.data(), sizeof(double)*n, cudaMemcpyHostToDevice));
    // exclusion masks by CALLER index: bit (d + 32) of atom i = "caller index i + d is excluded" (static)
    std::vector<unsigned long long> mask(n, 0ull);
    for (size_t k = 0; k + 1 < ctx->h_excl.size(); k += 2) {
        const int i = ctx->h_excl[k], j = ctx->h_excl[k+1];
        const int d = j - i;
        if (d >= -32 && d < 32) mask[i] |= 1ull << (d + 32);
        if (-d >= -32 && -d < 32) mask[j] |= 1ull << (-d + 32);
    }
    B2_CUDA(cudaMemcpy(D.exmask_user, mask.data(), sizeof(unsigned long long)*n, cudaMemcpyHostToDevice));
    D.nexcl = (int)ctx->h_excl.size()/2;
    if (D.nexcl > 0) {
        B2_CUDA(cudaMalloc(&D.excl_pairs, sizeof(int)*2*D.nexcl));
        B2_CUDA(cudaMemcpy(D.excl_pairs, ctx->h_excl.data(), sizeof(int)*2*D.nexcl, cudaMemcpyHostToDevice));
    }
    D.nsets = (int)ctx->h_sets.size();
    for (int k = 0; k < D.nsets; k++) {
        B2_CUDA(cudaMalloc(&D.sets_user[k], sizeof(double)*3*n));
        B2_CUDA(cudaMemcpy(D.sets_user[k], ctx->h_sets[k].data(), sizeof(double)*3*n, cudaMemcpyHostToDevice));
    }
    // sort / scan scratch: query both, keep the larger
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, D.keys, D.keys + nmol, D.ids, D.ids + nmol, nmol);
    cub::DeviceScan::ExclusiveSum(nullptr, b, D.sizes, D.offsets, nmol + 1);
    D.temp_bytes = std::max(a, b);
    B2_CUDA(cudaMalloc(&D.temp, std::max<size_t>(D.temp_bytes, 16)));
    D.valid = true;
    return B2_OK;
}

void order_release(b2_context* ctx) {
    OrderDevice& D = ctx->order;
    cudaFree(D.mol_ptr); cudaFree(D.mol_atoms); cudaFree(D.keys); cudaFree(D.ids); cudaFree(D.sizes); cudaFree(D.offsets);
    cudaFree(D.mass_user); cudaFree(D.exmask_user); cudaFree(D.span); cudaFree(D.excl_pairs); cudaFree(D.temp);
    for (int k = 0; k < B2_MAX_SETS; k++) cudaFree(D.sets_user[k]);
    D = OrderDevice();
}

// a parameter set changed on the host (updateParametersInContext): keep the caller-order device copy current
int order_refresh_param_set(b2_context* ctx, int k) {
    OrderDevice& D = ctx->order;
    if (!D.valid || k >= D.nsets || D.sets_user[k] == nullptr) return B2_OK;
    B2_CUDA(cudaMemcpy(D.sets_user[k], ctx->h_sets[k].data(), sizeof(double)*3*ctx->n, cudaMemcpyHostToDevice));
    return B2_OK;
}

// New spatial order from the configuration x_user (device, caller order): orig / inv / static tables on the device,
// h_orig on the host.
int order_compute_device(b2_context* ctx, const double* x_user) {
    B2_TRY(order_upload_tables(ctx));
    OrderDevice& D = ctx->order;
    const int n = ctx->n, nmol = D.nmol, T = 256;
    cudaStream_t s = ctx->stream;
    const int bits = order_hilbert_bits(ctx->box);
    k_mol_keys<<<(nmol + T - 1)/T, T, 0, s>>>(nmol, D.mol_ptr, D.mol_atoms, x_user, ctx->box[0], ctx->box[1], ctx->box[2], bits,
                                              D.keys, D.ids);
    B2_LAUNCH_CHECK();
    size_t bytes = D.temp_bytes;
    // stable LSD radix sort: equal keys keep the order of the molecule ids, like std::sort on (key, id) pairs
    if (cub::DeviceRadixSort::SortPairs(D.temp, bytes, D.keys, D.keys + nmol, D.ids, D.ids + nmol, nmol, 0, 64, s) != cudaSuccess)
        return b2_fail(ctx, B2_ERR_CUDA, "radix sort of the molecule keys failed: %s", cudaGetErrorString(cudaGetLastError()));
    const int* sorted = D.ids + nmol;
    k_mol_sizes<<<(nmol + T - 1)/T, T, 0, s>>>(nmol, sorted, D.mol_ptr, D.sizes);
    B2_LAUNCH_CHECK();
    bytes = D.temp_bytes;
    if (cub::DeviceScan::ExclusiveSum(D.temp, bytes, D.sizes, D.offsets, nmol + 1, s) != cudaSuccess)
        return b2_fail(ctx, B2_ERR_CUDA, "scan of the molecule sizes failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(ctx->mol_start);
    ctx->mol_start = nullptr;
    B2_CUDA(cudaMalloc(&ctx->mol_start, sizeof(int)*(D.nmol_used + 1)));
    ctx->nmol = D.nmol_used;
    k_fill_order<<<(nmol + T - 1)/T, T, 0, s>>>(nmol, n, D.nmol_used, sorted, D.offsets, D.mol_ptr, D.mol_atoms, ctx->orig, ctx->inv,
                                                ctx->mol_start);
    B2_LAUNCH_CHECK();
    GatherArgs g;
    memset(&g, 0, sizeof(g));
    g.nsets = D.nsets;
    for (int k = 0; k < D.nsets; k++) { g.sets_user[k] = D.sets_user[k]; g.par[k] = ctx->par[k]; g.pard[k] = ctx->pard[k]; }
    k_gather_static<<<(n + T - 1)/T, T, 0, s>>>(n, ctx->orig, D.mass_user, D.exmask_user, ctx->massd, ctx->invm, ctx->exmask, g);
    B2_LAUNCH_CHECK();
    B2_CUDA(cudaMemsetAsync(D.span, 0, sizeof(int), s));
    if (D.nexcl > 0) {
        k_excl_span<<<(D.nexcl + T - 1)/T, T, 0, s>>>(D.nexcl, D.excl_pairs, ctx->inv, D.span);
        B2_LAUNCH_CHECK();
    }
    ctx->h_orig.resize(n);
    B2_CUDA(cudaMemcpyAsync(ctx->h_orig.data(), ctx->orig, sizeof(int)*n, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaMemcpyAsync(&ctx->excl_span, D.span, sizeof(int), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    return B2_OK;
}

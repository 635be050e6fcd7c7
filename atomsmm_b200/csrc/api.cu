// C ABI of the engine (include/atomsmm_b200.h): context life cycle, system description, state
// transfer with spatial re-ordering, single-point evaluation and program loading.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <numeric>

#include "ctx.h"

#define PME_MIN_GRID 5
static thread_local std::string g_last_error;

int b2_fail(b2_context* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->err = buf;
    return code;
}

extern "C" const char* b2_last_error(const b2_context* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }
extern "C" const char* b2_version(void) { return "atomsmm_b200 0.1.0 (sm_100a)"; }

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_gather3(int n, const int* __restrict__ orig, const double* __restrict__ user, double* __restrict__ sorted) {
    int d = blockIdx.x*blockDim.x + threadIdx.x;
    if (d >= 3*n) return;
    sorted[d] = user[3*orig[d/3] + d%3];
}

__global__ void k_scatter3(int n, const int* __restrict__ orig, const double* __restrict__ sorted, double* __restrict__ user) {
    int d = blockIdx.x*blockDim.x + threadIdx.x;
    if (d >= 3*n) return;
    user[3*orig[d/3] + d%3] = sorted[d];
}

__global__ void k_scatter_force(int n, const int* __restrict__ orig, const float4* __restrict__ f, double* __restrict__ user) {
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = f[i];
    const int o = orig[i];
    user[3*o] = v.x; user[3*o+1] = v.y; user[3*o+2] = v.z;
}

__global__ void k_maxdisp(int n, const double* __restrict__ a, const double* __restrict__ b, double* out) {
    int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    double dx = a[3*i] - b[3*i], dy = a[3*i+1] - b[3*i+1], dz = a[3*i+2] - b[3*i+2];
    double d2 = dx*dx + dy*dy + dz*dz;
    // atoms that have strayed more than 0.5 nm from where they were when the order was computed
    if (d2 > 0.25) atomicAdd(reinterpret_cast<unsigned long long*>(out), 1ull);
}

__global__ void k_fold_energy(double* e, int group, double econst, int soft) {
    // scratch block at e+72: [0] energy, [1] virial, [2] dE/dlv, [3] dE/dlc
    e[group] += e[72] + econst;
    e[32 + group] += e[73];
    if (soft) { e[64] += e[74]; e[65] += e[75]; }
}

int state_permute_to_sorted(b2_context* ctx, const double* user, double* sorted) {
    const int T = 256;
    k_gather3<<<(3*ctx->n + T - 1)/T, T, 0, ctx->stream>>>(ctx->n, ctx->orig, user, sorted);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

int state_permute_to_user(b2_context* ctx, const double* sorted, double* user) {
    const int T = 256;
    k_scatter3<<<(3*ctx->n + T - 1)/T, T, 0, ctx->stream>>>(ctx->n, ctx->orig, sorted, user);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

// ---------------------------------------------------------------------------------------------
// life cycle
// ---------------------------------------------------------------------------------------------
extern "C" int b2_create(int device, b2_context** out) {
    b2_context* ctx = nullptr;
    if (!out) return b2_fail(nullptr, B2_ERR_ARG, "null output pointer");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return b2_fail(nullptr, B2_ERR_CUDA, "no CUDA device available (%s): this engine has no CPU path",
                       cudaGetErrorString(e));
    if (device < 0 || device >= count) return b2_fail(nullptr, B2_ERR_ARG, "device %d out of range", device);
    ctx = new b2_context();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return b2_fail(nullptr, B2_ERR_CUDA, "cannot initialise device %d", device);
    }
    ctx->own_stream = ctx->stream;
    if (cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return b2_fail(nullptr, B2_ERR_CUDA, "cannot create the side stream");
    }
    for (int g = 0; g < B2_FSLOTS; g++) ctx->fvalid[g] = -1;
    if (cudaMalloc(&ctx->d_energy, sizeof(double)*96) != cudaSuccess ||
        cudaMalloc(&ctx->rng_state, sizeof(unsigned long long)*4) != cudaSuccess ||
        cudaMalloc(&ctx->sum_partial, sizeof(double)*1024) != cudaSuccess ||
        cudaMalloc(&ctx->ticket, sizeof(unsigned)*4) != cudaSuccess ||
        cudaMalloc(&ctx->nl_flags, sizeof(int)*16) != cudaSuccess ||
        cudaMalloc(&ctx->band_count, sizeof(unsigned)*2) != cudaSuccess ||
        cudaMalloc(&ctx->band_ticket, sizeof(unsigned)*2) != cudaSuccess) {
        delete ctx;
        return b2_fail(nullptr, B2_ERR_CUDA, "device allocation failed");
    }
    cudaMemset(ctx->d_energy, 0, sizeof(double)*96);
    cudaMemset(ctx->rng_state, 0, sizeof(unsigned long long)*4);
    cudaMemset(ctx->ticket, 0, sizeof(unsigned)*4);
    cudaMemset(ctx->nl_flags, 0, sizeof(int)*16);
    cudaMemset(ctx->band_count, 0, sizeof(unsigned)*2);
    cudaMemset(ctx->band_ticket, 0, sizeof(unsigned)*2);
    ctx->sum_partial_size = 1024;
    *out = ctx;
    return B2_OK;
}

static void free_bonded(BondedForce& bf) {
    cudaFree(bf.atoms); cudaFree(bf.params); cudaFree(bf.code_e); cudaFree(bf.code_de); cudaFree(bf.consts);
}

extern "C" int b2_destroy(b2_context* ctx) {
    if (!ctx) return B2_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    program_release(ctx);
    dist_release(ctx);
    cudaFree(ctx->x); cudaFree(ctx->v); cudaFree(ctx->xref); cudaFree(ctx->xsort); cudaFree(ctx->xq);
    for (int s = 0; s < B2_MAX_SETS; s++) { cudaFree(ctx->par[s]); cudaFree(ctx->pard[s]); }
    cudaFree(ctx->massd); cudaFree(ctx->invm); cudaFree(ctx->orig); cudaFree(ctx->inv); cudaFree(ctx->exmask);
    order_release(ctx);
    jit_release(ctx);
    cudaFree(ctx->excl_ptr); cudaFree(ctx->excl_idx); cudaFree(ctx->scratch3);
    for (int g = 0; g < B2_FSLOTS; g++) cudaFree(ctx->fbuf[g]);
    for (double* p : ctx->perdof) cudaFree(p);
    for (int k = 0; k < B2_MAX_LISTS; k++) { cudaFree(ctx->lists[k].entries); cudaFree(ctx->lists[k].counts); cudaFree(ctx->lists[k].gflags); }
    cudaFree(ctx->cell_count); cudaFree(ctx->cell_start); cudaFree(ctx->cell_groups); cudaFree(ctx->gcell);
    cudaFree(ctx->gcen); cudaFree(ctx->ghalf); cudaFree(ctx->cgc); cudaFree(ctx->cgh); cudaFree(ctx->prel); cudaFree(ctx->fat_list);
    cudaFree(ctx->nl_flags); cudaFree(ctx->d_energy); cudaFree(ctx->code); cudaFree(ctx->consts);
    cudaFree(ctx->globals); cudaFree(ctx->sum_partial); cudaFree(ctx->rng_state);
    cudaFree(ctx->band_pairs); cudaFree(ctx->band_count); cudaFree(ctx->ticket);
    cudaFree(ctx->band_slot); cudaFree(ctx->band_acc); cudaFree(ctx->band_ticket);
    cudaFree(ctx->mol_start); cudaFree(ctx->xbackup); cudaFree(ctx->bond_acc);
    cudaFree(ctx->chunk_start); cudaFree(ctx->chunk_term_ptr); cudaFree(ctx->chunk_terms);
    for (double* p : ctx->carry_tmp) cudaFree(p);
    cudaFree(ctx->order_tmp);
    cudaFree(ctx->con_ptr); cudaFree(ctx->con_pairs); cudaFree(ctx->con_d2); cudaFree(ctx->xcon);
    for (BondedForce& bf : ctx->bonded_forces) free_bonded(bf);
    for (PmeForce& pm : ctx->pme_forces) pme_release(pm);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return B2_OK;
}

extern "C" int b2_set_stream(b2_context* ctx, void* cuda_stream) {
    if (!ctx) return B2_ERR_ARG;
    cudaStreamSynchronize(ctx->stream);
    program_release(ctx);
    ctx->stream = (cudaStream_t)cuda_stream;
    return B2_OK;
}

extern "C" int b2_synchronize(b2_context* ctx) {
    if (!ctx) return B2_ERR_ARG;
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->dd_state) {
        unsigned long long state[4];
        B2_CUDA(cudaMemcpy(state, ctx->dd_state, sizeof(state), cudaMemcpyDeviceToHost));
        if (state[2])
            return b2_fail(ctx, B2_ERR_STATE, "domain decomposition: a peer did not answer within the time-out "
                           "(exchange %llu); results since then are invalid", state[0]);
    }
    if (ctx->nl_flags) {
        int flags[11];
        B2_CUDA(cudaMemcpy(flags, ctx->nl_flags, sizeof(flags), cudaMemcpyDeviceToHost));
        if (flags[10])
            return b2_fail(ctx, B2_ERR_OVERFLOW, "cutoff-band buffer overflow (%u pairs): forces since then are incomplete",
                           ctx->band_capacity);
        ctx->counters[1] = flags[2];
        ctx->counters[4] = flags[3];
        if (flags[9])
            return b2_fail(ctx, B2_ERR_OVERFLOW, "SHAKE/RATTLE did not converge to the constraint tolerance %g", ctx->con_tol);
        if (flags[1])
            return b2_fail(ctx, B2_ERR_OVERFLOW, "neighbour-list capacity %d exceeded (largest list %d): results "
                           "since the last rebuild are invalid; set positions again to refit", ctx->lists[0].cap, flags[3]);
    }
    return B2_OK;
}

// ---------------------------------------------------------------------------------------------
// description
// ---------------------------------------------------------------------------------------------
extern "C" int b2_set_box(b2_context* ctx, const double box[3], int periodic) {
    if (!ctx || !box) return B2_ERR_ARG;
    for (int d = 0; d < 3; d++) {
        if (!(box[d] > 0)) return b2_fail(ctx, B2_ERR_ARG, "box length must be positive");
        ctx->box[d] = box[d];
    }
    ctx->periodic = periodic;
    ctx->lists_built = false;
    ctx->lists_fitted = false;
    ctx->have_order = false;
    program_release(ctx);
    return B2_OK;
}

extern "C" int b2_set_particles(b2_context* ctx, int n, const double* mass, const int* molecule) {
    if (!ctx || n <= 0 || !mass) return b2_fail(ctx, B2_ERR_ARG, "bad particle arguments");
    if (ctx->n != 0) return b2_fail(ctx, B2_ERR_STATE, "particles already set");
    if (n >= (1 << 24)) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "at most 2^24-1 atoms per context (list entries are 24-bit)");
    ctx->n = n;
    ctx->ngroups = (n + B2_GROUP - 1)/B2_GROUP;
    ctx->h_mass.assign(mass, mass + n);
    ctx->h_mol.resize(n);
    for (int i = 0; i < n; i++) ctx->h_mol[i] = molecule ? molecule[i] : i;
    B2_CUDA(cudaMalloc(&ctx->x, sizeof(double)*3*n));
    B2_CUDA(cudaMalloc(&ctx->v, sizeof(double)*3*n));
    B2_CUDA(cudaMalloc(&ctx->xref, sizeof(double)*3*n));
    B2_CUDA(cudaMalloc(&ctx->xsort, sizeof(double)*3*n));
    B2_CUDA(cudaMalloc(&ctx->xq, sizeof(int4)*n));
    B2_CUDA(cudaMalloc(&ctx->scratch3, sizeof(double)*3*n));
    B2_CUDA(cudaMalloc(&ctx->massd, sizeof(double)*n));
    B2_CUDA(cudaMalloc(&ctx->invm, sizeof(float)*n));
    B2_CUDA(cudaMalloc(&ctx->orig, sizeof(int)*n));
    B2_CUDA(cudaMalloc(&ctx->inv, sizeof(int)*n));
    B2_CUDA(cudaMalloc(&ctx->exmask, sizeof(unsigned long long)*n));
    B2_CUDA(cudaMemsetAsync(ctx->v, 0, sizeof(double)*3*n, ctx->stream));
    B2_CUDA(cudaMemsetAsync(ctx->exmask, 0, sizeof(unsigned long long)*n, ctx->stream));
    // cutoff-band buffers, per lane: a pair falls into the fp32 rounding band of the cutoff with probability
    // ~1e-5, i.e. ~2e-3 band pairs per atom at liquid densities; sized 16x above that
    if (ctx->band_capacity < (unsigned)(n/32)) ctx->band_capacity = (unsigned)(n/32);
    if (getenv("B2_BAND_CAPACITY")) ctx->band_capacity = (unsigned)std::max(1, atoi(getenv("B2_BAND_CAPACITY")));
    B2_CUDA(cudaMalloc(&ctx->band_pairs, sizeof(int)*4*(size_t)ctx->band_capacity));
    B2_CUDA(cudaMalloc(&ctx->band_acc, sizeof(long long)*6*(size_t)ctx->band_capacity));
    B2_CUDA(cudaMalloc(&ctx->band_slot, sizeof(int)*2*(size_t)n));
    B2_CUDA(cudaMemsetAsync(ctx->band_acc, 0, sizeof(long long)*6*(size_t)ctx->band_capacity, ctx->stream));
    B2_CUDA(cudaMemsetAsync(ctx->band_slot, 0xff, sizeof(int)*2*(size_t)n, ctx->stream));
    return B2_OK;
}

extern "C" int b2_add_param_set(b2_context* ctx, const double* charge, const double* sigma, const double* epsilon, int* set_id) {
    if (!ctx || ctx->n == 0) return b2_fail(ctx, B2_ERR_STATE, "set particles first");
    if ((int)ctx->h_sets.size() >= B2_MAX_SETS) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "too many parameter sets");
    std::vector<double> s(3*(size_t)ctx->n);
    for (int i = 0; i < ctx->n; i++) { s[3*i] = charge[i]; s[3*i+1] = sigma[i]; s[3*i+2] = epsilon[i]; }
    // reuse an identical set
    for (size_t k = 0; k < ctx->h_sets.size(); k++)
        if (ctx->h_sets[k] == s) { *set_id = (int)k; return B2_OK; }
    const int id = (int)ctx->h_sets.size();
    ctx->h_sets.push_back(s);
    B2_CUDA(cudaMalloc(&ctx->par[id], sizeof(float4)*ctx->n));
    B2_CUDA(cudaMalloc(&ctx->pard[id], sizeof(double)*3*ctx->n));
    ctx->have_order = false;
    *set_id = id;
    return B2_OK;
}

extern "C" int b2_set_exclusions(b2_context* ctx, int npairs, const int* pairs) {
    if (!ctx || ctx->n == 0) return b2_fail(ctx, B2_ERR_STATE, "set particles first");
    ctx->h_excl.assign(pairs, pairs + 2*(size_t)npairs);
    ctx->excl_far = false;
    for (int k = 0; k < npairs; k++) {
        int i = pairs[2*k], j = pairs[2*k+1];
        if (i < 0 || j < 0 || i >= ctx->n || j >= ctx->n) return b2_fail(ctx, B2_ERR_ARG, "exclusion index out of range");
        int d = j - i;
        if (d < -32 || d > 31 || -d < -32 || -d > 31) ctx->excl_far = true;
    }
    // CSR by caller index (used only when some exclusion spans more than 31 indices)
    std::vector<int> ptr(ctx->n + 1, 0);
    for (int k = 0; k < npairs; k++) { ptr[pairs[2*k] + 1]++; ptr[pairs[2*k+1] + 1]++; }
    for (int i = 0; i < ctx->n; i++) ptr[i+1] += ptr[i];
    std::vector<int> idx(std::max(1, 2*npairs)), cur(ptr.begin(), ptr.end() - 1);
    for (int k = 0; k < npairs; k++) {
        int i = pairs[2*k], j = pairs[2*k+1];
        idx[cur[i]++] = j; idx[cur[j]++] = i;
    }
    cudaFree(ctx->excl_ptr); cudaFree(ctx->excl_idx);
    B2_CUDA(cudaMalloc(&ctx->excl_ptr, sizeof(int)*(ctx->n + 1)));
    B2_CUDA(cudaMalloc(&ctx->excl_idx, sizeof(int)*idx.size()));
    B2_CUDA(cudaMemcpy(ctx->excl_ptr, ptr.data(), sizeof(int)*(ctx->n + 1), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(ctx->excl_idx, idx.data(), sizeof(int)*idx.size(), cudaMemcpyHostToDevice));
    ctx->have_order = false;
    ctx->order.valid = false;
    return B2_OK;
}

static int upload_param_set(b2_context* ctx, int k);

static int find_list(b2_context* ctx, double cutoff) {
    for (int k = 0; k < ctx->nlists; k++)
        if (fabs(ctx->lists[k].cutoff - cutoff) < 1e-12) return k;
    if (ctx->nlists >= B2_MAX_LISTS) return -1;
    ctx->lists[ctx->nlists].cutoff = cutoff;
    return ctx->nlists++;
}

extern "C" int b2_add_pair_force(b2_context* ctx, int family, int group, int param_set, double cutoff,
                                 const double* params, int nparams, double energy_constant, int* handle) {
    if (!ctx || ctx->n == 0) return b2_fail(ctx, B2_ERR_STATE, "set particles first");
    if (family < B2_PAIR_NEAR || family > B2_PAIR_SOFTCORE) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "unknown pair family %d", family);
    if (group < 0 || group > 31) return b2_fail(ctx, B2_ERR_ARG, "force group out of range");
    if (param_set < 0 || param_set >= (int)ctx->h_sets.size()) return b2_fail(ctx, B2_ERR_ARG, "unknown parameter set");
    if (nparams > B2_MAX_PAIR_PARAMS) return b2_fail(ctx, B2_ERR_ARG, "too many parameters");
    if (!(cutoff > 0)) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "pair forces need a finite cutoff");
    if (!ctx->periodic) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "pair forces need a periodic box");
    PairForce pf;
    memset(&pf, 0, sizeof(pf));
    pf.family = family; pf.group = group; pf.set = param_set; pf.cutoff = cutoff; pf.nparams = nparams;
    for (int k = 0; k < nparams; k++) pf.params[k] = params[k];
    pf.econst = energy_constant;
    for (int k = 0; k < B2_MAX_PAIR_PARAMS; k++) pf.bind[k] = -1;
    // the list radius only needs to cover the range of the potential
    double range = cutoff;
    if (family == B2_PAIR_NEAR || family == B2_PAIR_DAMPED) range = std::min(range, params[2]);
    pf.list = find_list(ctx, range);
    if (pf.list < 0) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "more than %d distinct cutoffs", B2_MAX_LISTS);
    ctx->pair_forces.push_back(pf);
    ctx->lists_built = false;
    ctx->lists_fitted = false;
    program_release(ctx);
    if (handle) *handle = (int)ctx->pair_forces.size() - 1;
    return B2_OK;
}

extern "C" int b2_update_pair_force(b2_context* ctx, int handle, const double* params, int nparams, double energy_constant) {
    if (!ctx || handle < 0 || handle >= (int)ctx->pair_forces.size()) return b2_fail(ctx, B2_ERR_ARG, "bad pair force handle");
    PairForce& pf = ctx->pair_forces[handle];
    if (nparams != pf.nparams) return b2_fail(ctx, B2_ERR_ARG, "parameter count mismatch");
    for (int k = 0; k < nparams; k++) pf.params[k] = params[k];
    pf.econst = energy_constant;
    for (int g = 0; g < B2_FSLOTS; g++) ctx->fvalid[g] = -1;
    ctx->deriv_version = -1;
    program_release(ctx);
    return B2_OK;
}

// Replaces the per-particle parameters of ONE pair force (updateParametersInContext after
// setParticleParameters; reference: utils.py reset_coulomb_scaling_factor-style rescaling).  A
// parameter set shared with other pair forces is split first, so that they keep their values.
extern "C" int b2_update_pair_particles(b2_context* ctx, int handle, const double* charge, const double* sigma,
                                        const double* epsilon) {
    if (!ctx || handle < 0 || handle >= (int)ctx->pair_forces.size()) return b2_fail(ctx, B2_ERR_ARG, "bad pair force handle");
    if (!charge || !sigma || !epsilon) return b2_fail(ctx, B2_ERR_ARG, "null parameter table");
    PairForce& pf = ctx->pair_forces[handle];
    const int n = ctx->n;
    std::vector<double> s(3*(size_t)n);
    for (int i = 0; i < n; i++) { s[3*i] = charge[i]; s[3*i+1] = sigma[i]; s[3*i+2] = epsilon[i]; }
    if (ctx->h_sets[pf.set] == s) return B2_OK;
    bool shared = false;
    for (size_t k = 0; k < ctx->pair_forces.size(); k++)
        if ((int)k != handle && ctx->pair_forces[k].set == pf.set) shared = true;
    for (const PmeForce& pm : ctx->pme_forces)
        if (pm.set == pf.set) shared = true;
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    int target = pf.set;
    if (shared) {
        if ((int)ctx->h_sets.size() >= B2_MAX_SETS) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "too many parameter sets");
        target = (int)ctx->h_sets.size();
        ctx->h_sets.push_back(s);
        B2_CUDA(cudaMalloc(&ctx->par[target], sizeof(float4)*n));
        B2_CUDA(cudaMalloc(&ctx->pard[target], sizeof(double)*3*n));
        pf.set = target;
    } else {
        ctx->h_sets[target] = s;
    }
    if (ctx->have_order) B2_TRY(upload_param_set(ctx, target));
    ctx->order.valid = false;             // the caller-order copy the device ordering gathers from is stale
    for (int g = 0; g < B2_FSLOTS; g++) ctx->fvalid[g] = -1;
    ctx->deriv_version = -1;
    program_release(ctx);
    return B2_OK;
}

extern "C" int b2_bind_pair_parameter(b2_context* ctx, int handle, int param, int global_index) {
    if (!ctx || handle < 0 || handle >= (int)ctx->pair_forces.size()) return b2_fail(ctx, B2_ERR_ARG, "bad pair force handle");
    PairForce& pf = ctx->pair_forces[handle];
    if (pf.family != B2_PAIR_SOFTCORE || (param != 1 && param != 2))
        return b2_fail(ctx, B2_ERR_UNSUPPORTED, "only the soft-core couplings (parameters 1, 2) can be bound to integrator globals");
    if (global_index >= ctx->nglobals) return b2_fail(ctx, B2_ERR_ARG, "global index out of range");
    pf.bind[param] = global_index;
    for (int g = 0; g < B2_FSLOTS; g++) ctx->fvalid[g] = -1;
    ctx->deriv_version = -1;
    program_release(ctx);
    return B2_OK;
}

static int add_bonded_common(b2_context* ctx, BondedForce& bf, const int* atoms, const double* params) {
    const int arity = bf.arity;
    for (int k = 0; k < arity*bf.nterms; k++)
        if (atoms[k] < 0 || atoms[k] >= ctx->n) return b2_fail(ctx, B2_ERR_ARG, "bonded atom index out of range");
    if (bf.nterms > 0) {
        bf.h_atoms.assign(atoms, atoms + (size_t)arity*bf.nterms);
        B2_CUDA(cudaMalloc(&bf.atoms, sizeof(int)*arity*bf.nterms));
        B2_CUDA(cudaMemcpy(bf.atoms, atoms, sizeof(int)*arity*bf.nterms, cudaMemcpyHostToDevice));
        const size_t np = (size_t)std::max(1, bf.stride)*bf.nterms;
        B2_CUDA(cudaMalloc(&bf.params, sizeof(double)*np));
        if (bf.stride > 0) B2_CUDA(cudaMemcpy(bf.params, params, sizeof(double)*np, cudaMemcpyHostToDevice));
    }
    return B2_OK;
}

extern "C" int b2_add_bonded_force(b2_context* ctx, int family, int group, int nterms, const int* atoms,
                                   const double* params, int stride, int periodic, const double* gparams,
                                   int ngparams, int* handle) {
    if (!ctx || ctx->n == 0) return b2_fail(ctx, B2_ERR_STATE, "set particles first");
    BondedForce bf;
    bf.family = family; bf.group = group; bf.nterms = nterms; bf.stride = stride; bf.periodic = periodic;
    switch (family) {
    case B2_BOND_HARMONIC: case B2_BOND_LJC: bf.arity = 2; break;
    case B2_ANGLE_HARMONIC: bf.arity = 3; break;
    case B2_TORSION_PERIODIC: bf.arity = 4; break;
    default: return b2_fail(ctx, B2_ERR_UNSUPPORTED, "unknown bonded family %d", family);
    }
    if (ngparams > 8) return b2_fail(ctx, B2_ERR_ARG, "too many global parameters");
    for (int k = 0; k < ngparams; k++) bf.gparams[k] = gparams[k];
    B2_TRY(add_bonded_common(ctx, bf, atoms, params));
    ctx->bonded_forces.push_back(bf);
    ctx->inner_built = false;
    program_release(ctx);
    if (handle) *handle = (int)ctx->bonded_forces.size() - 1;
    return B2_OK;
}

extern "C" int b2_add_custom_bonded_force(b2_context* ctx, int family, int group, int nterms, const int* atoms,
                                          const double* params, int stride, int periodic, const int* code_e,
                                          int ncode_e, const int* code_de, int ncode_de, const double* consts,
                                          int nconsts, int* handle) {
    if (!ctx || ctx->n == 0) return b2_fail(ctx, B2_ERR_STATE, "set particles first");
    if (family != B2_BOND_CUSTOM && family != B2_ANGLE_CUSTOM) return b2_fail(ctx, B2_ERR_ARG, "not a custom family");
    if (stride > 9) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "at most 9 per-term parameters");
    BondedForce bf;
    bf.family = family; bf.group = group; bf.nterms = nterms; bf.stride = stride; bf.periodic = periodic;
    bf.arity = family == B2_BOND_CUSTOM ? 2 : 3;
    B2_TRY(add_bonded_common(ctx, bf, atoms, params));
    bf.ncode_e = ncode_e; bf.ncode_de = ncode_de;
    B2_CUDA(cudaMalloc(&bf.code_e, sizeof(int)*2*std::max(1, ncode_e)));
    B2_CUDA(cudaMalloc(&bf.code_de, sizeof(int)*2*std::max(1, ncode_de)));
    B2_CUDA(cudaMalloc(&bf.consts, sizeof(double)*std::max(1, nconsts)));
    if (ncode_e) B2_CUDA(cudaMemcpy(bf.code_e, code_e, sizeof(int)*2*ncode_e, cudaMemcpyHostToDevice));
    if (ncode_de) B2_CUDA(cudaMemcpy(bf.code_de, code_de, sizeof(int)*2*ncode_de, cudaMemcpyHostToDevice));
    if (nconsts) B2_CUDA(cudaMemcpy(bf.consts, consts, sizeof(double)*nconsts, cudaMemcpyHostToDevice));
    ctx->bonded_forces.push_back(bf);
    ctx->inner_built = false;
    program_release(ctx);
    if (handle) *handle = (int)ctx->bonded_forces.size() - 1;
    return B2_OK;
}

extern "C" int b2_add_pme(b2_context* ctx, int group, int param_set, double alpha, int nx, int ny, int nz, double kc,
                          double self_energy, int* handle) {
    if (!ctx || ctx->n == 0) return b2_fail(ctx, B2_ERR_STATE, "set particles first");
    if (param_set < 0 || param_set >= (int)ctx->h_sets.size()) return b2_fail(ctx, B2_ERR_ARG, "unknown parameter set");
    if (nx < PME_MIN_GRID || ny < PME_MIN_GRID || nz < PME_MIN_GRID || !(alpha > 0))
        return b2_fail(ctx, B2_ERR_ARG, "bad PME parameters");
    if (!ctx->periodic) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "PME needs a periodic box");
    if (ctx->nranks > 1) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "PME is not available with domain decomposition yet");
    PmeForce pf;
    pf.group = group; pf.set = param_set; pf.alpha = alpha; pf.kc = kc; pf.eself = self_energy;
    pf.K[0] = nx; pf.K[1] = ny; pf.K[2] = nz;
    B2_TRY(pme_setup(ctx, pf));
    ctx->pme_forces.push_back(pf);
    program_release(ctx);
    if (handle) *handle = (int)ctx->pme_forces.size() - 1;
    return B2_OK;
}

extern "C" int b2_set_skin(b2_context* ctx, double skin) {
    if (!ctx || !(skin >= 0)) return B2_ERR_ARG;
    ctx->skin = skin;
    ctx->lists_built = false;
    ctx->lists_fitted = false;
    ctx->have_order = false;
    return B2_OK;
}

// ---------------------------------------------------------------------------------------------
// spatial ordering (host side; runs when positions are set for the first time or have moved far
// from the configuration the current order was computed for)
// ---------------------------------------------------------------------------------------------
// Hilbert-curve index of a cell (Skilling's transpose algorithm).  Unlike the Morton (Z) curve
// the Hilbert curve has no jumps: consecutive keys are always face-adjacent cells, so every run of
// consecutive molecules -- in particular every 8-atom i-group -- is spatially compact.
static uint64_t hilbert_key(const uint32_t cell[3], int bits) {
    uint32_t X[3] = {cell[0], cell[1], cell[2]};
    const uint32_t M = 1u << (bits - 1);
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
        for (int i = 0; i < 3; i++) {
            if (X[i] & Q) X[0] ^= P;
            else { const uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    for (int i = 1; i < 3; i++) X[i] ^= X[i-1];
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1;
    for (int i = 0; i < 3; i++) X[i] ^= t;
    uint64_t key = 0;
    for (int q = bits - 1; q >= 0; q--)
        for (int i = 0; i < 3; i++) key = (key << 1) | ((X[i] >> q) & 1u);
    return key;
}

// Pure host helper (no GPU needed): the Hilbert index the engine sorts molecules by, for a point in a
// periodic box; exported so that the ordering can be tested on the CPU.
extern "C" int b2_hilbert_index(const double position[3], const double box[3], unsigned long long* out_key) {
    if (!position || !box || !out_key) return B2_ERR_ARG;
    const double longest = std::max(box[0], std::max(box[1], box[2]));
    int bits = 1;
    while (bits < 20 && longest/(double)(1u << bits) > 0.32) bits++;
    uint32_t c[3];
    for (int d = 0; d < 3; d++) {
        const double L = box[d];
        if (!(L > 0)) return B2_ERR_ARG;
        const double w = position[d] - L*floor(position[d]/L);
        const int k = (int)(w/L*(double)(1u << bits));
        c[d] = (uint32_t)std::min(std::max(k, 0), (1 << bits) - 1);
    }
    *out_key = hilbert_key(c, bits);
    return B2_OK;
}

extern "C" int b2_set_jit(b2_context* ctx, int enabled) {
    if (!ctx) return B2_ERR_ARG;
    if (ctx->jit_enabled != (enabled != 0)) {
        program_release(ctx);
        jit_release(ctx);
        ctx->jit_enabled = enabled != 0;
    }
    return B2_OK;
}

extern "C" int b2_get_jit_stats(b2_context* ctx, long long out_host[2]) {
    if (!ctx || !out_host) return B2_ERR_ARG;
    out_host[0] = ctx->jit_compiled;
    out_host[1] = ctx->jit_launches;
    return B2_OK;
}

extern "C" int b2_get_order(b2_context* ctx, int* orig_host) {
    if (!ctx || !orig_host) return B2_ERR_ARG;
    if (!ctx->have_order) return b2_fail(ctx, B2_ERR_STATE, "positions have not been set");
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    B2_CUDA(cudaMemcpy(orig_host, ctx->orig, sizeof(int)*ctx->n, cudaMemcpyDeviceToHost));
    return B2_OK;
}

static int compute_order(b2_context* ctx, const std::vector<double>& hx) {
    const int n = ctx->n;
    // molecules in caller order: a static CSR table, built once (config 5 has 1.4 M molecules; the order is
    // refreshed during long runs, so this path must not allocate per molecule)
    if (ctx->h_mol_ptr.empty()) {
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
    }
    const int nmol = (int)ctx->h_mol_ptr.size() - 1;
    std::vector<std::pair<uint64_t, int>> keys;
    keys.reserve(nmol);
    for (int m = 0; m < nmol; m++) {
        if (ctx->h_mol_ptr[m+1] == ctx->h_mol_ptr[m]) continue;
        unsigned long long key = 0;
        b2_hilbert_index(&hx[3*(size_t)ctx->h_mol_atoms[ctx->h_mol_ptr[m]]], ctx->box, &key);   // by the molecule's first atom
        keys.emplace_back((uint64_t)key, m);
    }
    std::sort(keys.begin(), keys.end());        // pairs (key, molecule id): ties resolved by the id, i.e. stable
    ctx->h_orig.clear();
    ctx->h_orig.reserve(n);
    for (auto& km : keys)
        for (int k = ctx->h_mol_ptr[km.second]; k < ctx->h_mol_ptr[km.second + 1]; k++) ctx->h_orig.push_back(ctx->h_mol_atoms[k]);
    return B2_OK;
}

// parameter set k in the engine's order: fp32 {q, sigma/2, sqrt(eps)} for the force tiles, fp64 {q, sigma, eps}
static int upload_param_set(b2_context* ctx, int k) {
    const int n = ctx->n;
    std::vector<float4> p(n);
    std::vector<double> pd(3*(size_t)n);
    for (int s = 0; s < n; s++) {
        const double* src = &ctx->h_sets[k][3*(size_t)ctx->h_orig[s]];
        p[s] = make_float4((float)src[0], (float)(0.5*src[1]), (float)sqrt(src[2]), 0.f);
        pd[3*s] = src[0]; pd[3*s+1] = src[1]; pd[3*s+2] = src[2];
    }
    B2_CUDA(cudaMemcpy(ctx->par[k], p.data(), sizeof(float4)*n, cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(ctx->pard[k], pd.data(), sizeof(double)*3*n, cudaMemcpyHostToDevice));
    return B2_OK;
}

static int upload_static(b2_context* ctx) {
    const int n = ctx->n;
    std::vector<int> inv(n);
    for (int s = 0; s < n; s++) inv[ctx->h_orig[s]] = s;
    B2_CUDA(cudaMemcpy(ctx->orig, ctx->h_orig.data(), sizeof(int)*n, cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(ctx->inv, inv.data(), sizeof(int)*n, cudaMemcpyHostToDevice));
    std::vector<double> md(n);
    std::vector<float> im(n);
    for (int s = 0; s < n; s++) {
        const double m = ctx->h_mass[ctx->h_orig[s]];
        md[s] = m;
        im[s] = m > 0 ? (float)(1.0/m) : 0.f;
    }
    B2_CUDA(cudaMemcpy(ctx->massd, md.data(), sizeof(double)*n, cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(ctx->invm, im.data(), sizeof(float)*n, cudaMemcpyHostToDevice));
    for (size_t k = 0; k < ctx->h_sets.size(); k++) B2_TRY(upload_param_set(ctx, (int)k));
    // molecule table in the engine's order (molecules are contiguous): barostat moves, chunking
    std::vector<int> mstart;
    for (int s = 0; s < n; s++)
        if (s == 0 || ctx->h_mol[ctx->h_orig[s]] != ctx->h_mol[ctx->h_orig[s-1]]) mstart.push_back(s);
    ctx->nmol = (int)mstart.size();
    mstart.push_back(n);
    cudaFree(ctx->mol_start);
    ctx->mol_start = nullptr;
    B2_CUDA(cudaMalloc(&ctx->mol_start, sizeof(int)*mstart.size()));
    B2_CUDA(cudaMemcpy(ctx->mol_start, mstart.data(), sizeof(int)*mstart.size(), cudaMemcpyHostToDevice));
    std::vector<unsigned long long> mask(n, 0ull);   // indexed by caller index first
    ctx->excl_span = 0;
    for (size_t k = 0; k + 1 < ctx->h_excl.size(); k += 2) {
        const int i = ctx->h_excl[k], j = ctx->h_excl[k+1];
        ctx->excl_span = std::max(ctx->excl_span, abs(inv[i] - inv[j]));
        const int d = j - i;
        if (d >= -32 && d < 32) mask[i] |= 1ull << (d + 32);
        if (-d >= -32 && -d < 32) mask[j] |= 1ull << (-d + 32);
    }
    std::vector<unsigned long long> ms(n);
    for (int s = 0; s < n; s++) ms[s] = mask[ctx->h_orig[s]];
    B2_CUDA(cudaMemcpy(ctx->exmask, ms.data(), sizeof(unsigned long long)*n, cudaMemcpyHostToDevice));
    return B2_OK;
}

extern "C" int b2_set_positions(b2_context* ctx, const double* x_dev) {
    if (!ctx || ctx->n == 0 || !x_dev) return b2_fail(ctx, B2_ERR_STATE, "set particles first");
    const int n = ctx->n, T = 256;
    B2_TRY(dist_before_move(ctx));
    bool resort = !ctx->have_order || ctx->force_resort;
    ctx->force_resort = false;
    if (!resort) {
        // has the configuration drifted away from the one the order was built for?
        B2_TRY(state_permute_to_sorted(ctx, x_dev, ctx->scratch3));
        double* flag = ctx->d_energy + 80;
        B2_CUDA(cudaMemsetAsync(flag, 0, sizeof(double), ctx->stream));
        k_maxdisp<<<(n + T - 1)/T, T, 0, ctx->stream>>>(n, ctx->scratch3, ctx->xsort, flag);
        B2_LAUNCH_CHECK();
        unsigned long long strayed = 0;
        B2_CUDA(cudaMemcpyAsync(&strayed, flag, sizeof(strayed), cudaMemcpyDeviceToHost, ctx->stream));
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
        // the order only matters for speed (compact groups, short lists): refresh it when more than 2 %
        // of the atoms have strayed, not for the first fast hydrogen
        resort = strayed*50ull > (unsigned long long)n;
    }
    static const bool timing = getenv("B2_DEBUG_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_mark = now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        cudaStreamSynchronize(ctx->stream);
        const double t = now();
        fprintf(stderr, "[b2 resort] %-18s %8.2f ms\n", what, t - t_mark);
        t_mark = t;
    };
    if (resort) {
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
        // carry velocities and per-DOF variables across the re-ordering (via caller order)
        std::vector<double*> carried;
        if (ctx->have_order) {
            carried.push_back(ctx->v);
            for (double* p : ctx->perdof) carried.push_back(p);
            // with several ranks each holds only its owned range: complete the arrays first, because
            // the ownership ranges change with the order
            for (double* p : carried) B2_TRY(dist_gather3(ctx, p));
        }
        // staging buffers are kept for the life of the context: no allocation on this path
        std::vector<double*>& tmp = ctx->carry_tmp;
        while (tmp.size() < carried.size()) {
            double* p = nullptr;
            B2_CUDA(cudaMalloc(&p, sizeof(double)*3*n));
            tmp.push_back(p);
        }
        for (size_t k = 0; k < carried.size(); k++) B2_TRY(state_permute_to_user(ctx, carried[k], tmp[k]));
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
        lap("download+carry");
        static const bool host_order = getenv("B2_HOST_ORDER") != nullptr;
        if (host_order) {           // round 1's path, kept for A/B checks of the device ordering
            std::vector<double> hx(3*(size_t)n);
            B2_CUDA(cudaMemcpy(hx.data(), x_dev, sizeof(double)*3*n, cudaMemcpyDeviceToHost));
            B2_TRY(compute_order(ctx, hx));
            lap("compute_order");
            B2_TRY(upload_static(ctx));
            lap("upload_static");
        } else {
            B2_TRY(order_compute_device(ctx, x_dev));
            lap("device order");
        }
        B2_TRY(dist_partition(ctx));
        ctx->inner_built = false;
        ctx->con_built = false;
        for (size_t k = 0; k < carried.size(); k++) {
            B2_TRY(state_permute_to_sorted(ctx, tmp[k], carried[k]));
        }
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->have_order = true;
        B2_TRY(state_permute_to_sorted(ctx, x_dev, ctx->x));
        B2_CUDA(cudaMemcpyAsync(ctx->xsort, ctx->x, sizeof(double)*3*n, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->lists_built = false;
        program_release(ctx);
        lap("permute+release");
    } else {
        B2_CUDA(cudaMemcpyAsync(ctx->x, ctx->scratch3, sizeof(double)*3*n, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    ctx->have_positions = true;
    ctx->pos_version++;
    ctx->x_synced = ctx->pos_version;     // the caller passes the full configuration on every rank
    if (!ctx->lists_built && ctx->nlists > 0) {
        B2_TRY(nl_initial_build(ctx));
        lap("list build");
    }
    return B2_OK;
}

// Called between MD steps of a long run: molecules diffuse, so the spatial order the groups, chunks
// and ownership ranges were built for decays (lists stay exact but grow).  When enough atoms have
// strayed, go through the re-ordering path of b2_set_positions with the current configuration.
int order_refresh(b2_context* ctx) {
    if (!ctx->have_order || !ctx->have_positions) return B2_OK;
    const int n = ctx->n, T = 256;
    B2_TRY(dist_sync_positions(ctx));
    double* flag = ctx->d_energy + 80;
    B2_CUDA(cudaMemsetAsync(flag, 0, sizeof(double), ctx->stream));
    k_maxdisp<<<(n + T - 1)/T, T, 0, ctx->stream>>>(n, ctx->x, ctx->xsort, flag);
    B2_LAUNCH_CHECK();
    unsigned long long strayed = 0;
    B2_CUDA(cudaMemcpyAsync(&strayed, flag, sizeof(strayed), cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    if (strayed*50ull <= (unsigned long long)n) return B2_OK;
    if (ctx->order_tmp == nullptr) B2_CUDA(cudaMalloc(&ctx->order_tmp, sizeof(double)*3*n));
    double* user = ctx->order_tmp;
    int r = state_permute_to_user(ctx, ctx->x, user);
    if (r == B2_OK) {
        ctx->force_resort = true;
        r = b2_set_positions(ctx, user);
    }
    cudaStreamSynchronize(ctx->stream);
    ctx->counters[7] += 1000000;       // visible in b2_comm_info / counters: re-orderings during runs x 1e6
    return r;
}

extern "C" int b2_set_velocities(b2_context* ctx, const double* v_dev) {
    if (!ctx || !ctx->have_order) return b2_fail(ctx, B2_ERR_STATE, "set positions before velocities");
    ctx->v_version++;
    return state_permute_to_sorted(ctx, v_dev, ctx->v);
}

extern "C" int b2_get_positions(b2_context* ctx, double* x_dev) {
    if (!ctx || !ctx->have_positions) return b2_fail(ctx, B2_ERR_STATE, "positions have not been set");
    B2_TRY(dist_sync_positions(ctx));
    return state_permute_to_user(ctx, ctx->x, x_dev);
}

extern "C" int b2_get_velocities(b2_context* ctx, double* v_dev) {
    if (!ctx || !ctx->have_order) return b2_fail(ctx, B2_ERR_STATE, "positions have not been set");
    B2_TRY(dist_gather3(ctx, ctx->v));
    return state_permute_to_user(ctx, ctx->v, v_dev);
}

// ---------------------------------------------------------------------------------------------
// single-point evaluation
// ---------------------------------------------------------------------------------------------
extern "C" int b2_eval(b2_context* ctx, uint32_t group_mask, int flags, double* forces_dev, double* energy_host,
                       double* virial_host) {
    if (!ctx || !ctx->have_positions) return b2_fail(ctx, B2_ERR_STATE, "positions have not been set");
    const int n = ctx->n, T = 256;
    if ((flags & B2_EVAL_FORCES) && (flags & B2_EVAL_DOUBLE) && forces_dev) {
        // report cadence: every contribution evaluated and accumulated in float64 (engine order in scratch3)
        double* f64 = ctx->scratch3;
        B2_CUDA(cudaMemsetAsync(f64, 0, sizeof(double)*3*n, ctx->stream));
        bool any_pair = false;
        for (const PairForce& pf : ctx->pair_forces)
            if (group_mask & (1u << pf.group)) any_pair = true;
        if (any_pair) {
            if (!ctx->p2p) B2_TRY(dist_sync_positions(ctx));
            B2_TRY(nl_prepare(ctx, false));
        }
        for (const PairForce& pf : ctx->pair_forces)
            if (group_mask & (1u << pf.group)) B2_TRY(pair_eval_forces64(ctx, pf, f64));
        B2_TRY(dist_before_move(ctx));
        B2_TRY(bonded_eval_forces64(ctx, group_mask, f64));
        for (PmeForce& pm : ctx->pme_forces)
            if (group_mask & (1u << pm.group)) B2_TRY(pme_eval(ctx, pm, nullptr, nullptr, f64));
        B2_TRY(dist_gather3(ctx, f64));
        B2_TRY(state_permute_to_user(ctx, f64, forces_dev));
    } else if ((flags & B2_EVAL_FORCES) && forces_dev) {
        int slot = 32;
        bool scratch = true;
        if (group_mask == 0xffffffffu) scratch = false;
        else if (__builtin_popcount(group_mask) == 1) { slot = __builtin_ctz(group_mask); scratch = false; }
        if (scratch) ctx->fvalid[32] = -1;
        B2_TRY(forces_ensure(ctx, group_mask, slot));
        if (scratch) ctx->fvalid[32] = -1;
        B2_TRY(dist_gather_forces(ctx, ctx->fbuf[slot]));
        k_scatter_force<<<(n + T - 1)/T, T, 0, ctx->stream>>>(n, ctx->orig, ctx->fbuf[slot], forces_dev);
        B2_LAUNCH_CHECK();
    }
    if (flags & B2_EVAL_ENERGY) {
        B2_CUDA(cudaMemsetAsync(ctx->d_energy, 0, sizeof(double)*72, ctx->stream));
        bool any_pair = false;
        for (const PairForce& pf : ctx->pair_forces)
            if (group_mask & (1u << pf.group)) any_pair = true;
        if (any_pair) {
            if (!ctx->p2p) B2_TRY(dist_sync_positions(ctx));
            B2_TRY(nl_prepare(ctx, false));
        }
        for (const PairForce& pf : ctx->pair_forces) {
            if (!(group_mask & (1u << pf.group))) continue;
            B2_TRY(pair_eval_energy(ctx, pf, pf.group));
            B2_TRY(dist_allreduce(ctx, ctx->d_energy + 72, 4));
            k_fold_energy<<<1, 1, 0, ctx->stream>>>(ctx->d_energy, pf.group, pf.econst, pf.family == B2_PAIR_SOFTCORE);
            B2_LAUNCH_CHECK();
        }
        for (const BondedForce& bf : ctx->bonded_forces) {
            if (!(group_mask & (1u << bf.group)) || bf.nterms == 0) continue;
            B2_TRY(bonded_eval(ctx, bf, nullptr, false, true));
            B2_TRY(dist_allreduce(ctx, ctx->d_energy + 72, 4));
            k_fold_energy<<<1, 1, 0, ctx->stream>>>(ctx->d_energy, bf.group, 0.0, 0);
            B2_LAUNCH_CHECK();
        }
        for (PmeForce& pm : ctx->pme_forces) {
            if (!(group_mask & (1u << pm.group))) continue;
            B2_CUDA(cudaMemsetAsync(ctx->d_energy + 72, 0, 4*sizeof(double), ctx->stream));
            B2_TRY(pme_eval(ctx, pm, nullptr, ctx->d_energy + 72));
            k_fold_energy<<<1, 1, 0, ctx->stream>>>(ctx->d_energy, pm.group, pm.eself, 0);
            B2_LAUNCH_CHECK();
        }
        B2_TRY(dist_before_move(ctx));
        double h[66];
        B2_CUDA(cudaMemcpyAsync(h, ctx->d_energy, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        B2_TRY(b2_synchronize(ctx));
        double e = 0, w = 0;
        for (int g = 0; g < 32; g++) {
            ctx->h_energy[g] = h[g]; ctx->h_virial[g] = h[32+g];
            if (group_mask & (1u << g)) { e += h[g]; w += h[32+g]; }
        }
        ctx->h_dlambda[0] = h[64]; ctx->h_dlambda[1] = h[65];
        if (energy_host) *energy_host = e;
        if (virial_host) *virial_host = w;
    }
    return B2_OK;
}

extern "C" int b2_get_group_energies(b2_context* ctx, double energy_host[32], double virial_host[32]) {
    if (!ctx) return B2_ERR_ARG;
    for (int g = 0; g < 32; g++) {
        if (energy_host) energy_host[g] = ctx->h_energy[g];
        if (virial_host) virial_host[g] = ctx->h_virial[g];
    }
    return B2_OK;
}

extern "C" int b2_get_parameter_derivatives(b2_context* ctx, double out_host[2]) {
    if (!ctx) return B2_ERR_ARG;
    out_host[0] = ctx->h_dlambda[0]; out_host[1] = ctx->h_dlambda[1];
    return B2_OK;
}

extern "C" int b2_pair_set(b2_context* ctx, int handle, long long* count_host, unsigned long long* checksum_host,
                           int* pairs_dev, long long capacity) {
    if (!ctx || handle < 0 || handle >= (int)ctx->pair_forces.size()) return b2_fail(ctx, B2_ERR_ARG, "bad pair force handle");
    if (!ctx->have_positions) return b2_fail(ctx, B2_ERR_STATE, "positions have not been set");
    B2_TRY(dist_sync_positions(ctx));
    B2_TRY(nl_prepare(ctx, false));
    return pair_count_set(ctx, ctx->pair_forces[handle], count_host, checksum_host, pairs_dev, capacity);
}

// ---------------------------------------------------------------------------------------------
// program
// ---------------------------------------------------------------------------------------------
extern "C" int b2_load_program(b2_context* ctx, const int* ops, int nops, const int* code, int ncode,
                               const double* consts, int nconsts, const double* globals, int nglobals,
                               int nperdof, uint64_t seed) {
    if (!ctx || ctx->n == 0) return b2_fail(ctx, B2_ERR_STATE, "set particles first");
    if (nperdof > B2_MAX_PERDOF) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "at most %d per-DOF variables", B2_MAX_PERDOF);
    program_release(ctx);
    jit_release(ctx);
    for (PairForce& pf : ctx->pair_forces)
        for (int k = 0; k < B2_MAX_PAIR_PARAMS; k++) pf.bind[k] = -1;
    ctx->ops.resize(nops);
    for (int k = 0; k < nops; k++) {
        const int* w = ops + (size_t)k*B2_OP_WORDS;
        ctx->ops[k] = b2_op{w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]};
    }
    cudaFree(ctx->code); cudaFree(ctx->consts); cudaFree(ctx->globals);
    ctx->code = nullptr; ctx->consts = nullptr; ctx->globals = nullptr;
    ctx->ncode = ncode; ctx->nconsts = nconsts; ctx->nglobals = nglobals;
    B2_CUDA(cudaMalloc(&ctx->code, sizeof(int)*std::max(2, ncode)));
    B2_CUDA(cudaMalloc(&ctx->consts, sizeof(double)*std::max(1, nconsts)));
    B2_CUDA(cudaMalloc(&ctx->globals, sizeof(double)*std::max(1, nglobals)));
    ctx->h_code.assign(code, code + ncode);
    if (ncode) B2_CUDA(cudaMemcpy(ctx->code, code, sizeof(int)*ncode, cudaMemcpyHostToDevice));
    if (nconsts) B2_CUDA(cudaMemcpy(ctx->consts, consts, sizeof(double)*nconsts, cudaMemcpyHostToDevice));
    if (nglobals) B2_CUDA(cudaMemcpy(ctx->globals, globals, sizeof(double)*nglobals, cudaMemcpyHostToDevice));
    while ((int)ctx->perdof.size() < nperdof) {
        double* p = nullptr;
        B2_CUDA(cudaMalloc(&p, sizeof(double)*3*ctx->n));
        B2_CUDA(cudaMemset(p, 0, sizeof(double)*3*ctx->n));
        ctx->perdof.push_back(p);
    }
    unsigned long long st[4] = {seed, 0ull, 0ull, 0ull};
    B2_CUDA(cudaMemcpy(ctx->rng_state, st, sizeof(st), cudaMemcpyHostToDevice));
    ctx->program_loaded = true;
    ctx->mvv_index = -1; ctx->mvv_version = -1;
    ctx->eager_steps = 0;
    ctx->prologue_valid = false;
    ctx->uses_random = false;
    for (int k = 0; k + 1 < ncode; k += 2)
        if (code[k] == VM_GAUSS || code[k] == VM_UNIF) ctx->uses_random = true;
    return B2_OK;
}

extern "C" int b2_set_globals(b2_context* ctx, int first, int count, const double* values_host) {
    if (!ctx || first < 0 || first + count > ctx->nglobals) return b2_fail(ctx, B2_ERR_ARG, "global index out of range");
    B2_CUDA(cudaMemcpyAsync(ctx->globals + first, values_host, sizeof(double)*count, cudaMemcpyHostToDevice, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->mvv_version = -1;               // the caller may have overwritten the carried sum(m v.v)
    if (ctx->prologue_valid) {           // coefficients derived from globals must be recomputed
        ctx->prologue_valid = false;
        program_release(ctx);
    }
    return B2_OK;
}

extern "C" int b2_get_globals(b2_context* ctx, int first, int count, double* values_host) {
    if (!ctx || first < 0 || first + count > ctx->nglobals) return b2_fail(ctx, B2_ERR_ARG, "global index out of range");
    B2_CUDA(cudaMemcpyAsync(values_host, ctx->globals + first, sizeof(double)*count, cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    return B2_OK;
}

extern "C" int b2_set_perdof(b2_context* ctx, int var, const double* values_dev) {
    if (!ctx || var < 0 || var >= (int)ctx->perdof.size()) return b2_fail(ctx, B2_ERR_ARG, "per-DOF variable out of range");
    if (!ctx->have_order) return b2_fail(ctx, B2_ERR_STATE, "set positions first");
    return state_permute_to_sorted(ctx, values_dev, ctx->perdof[var]);
}

extern "C" int b2_get_perdof(b2_context* ctx, int var, double* values_dev) {
    if (!ctx || var < 0 || var >= (int)ctx->perdof.size()) return b2_fail(ctx, B2_ERR_ARG, "per-DOF variable out of range");
    if (!ctx->have_order) return b2_fail(ctx, B2_ERR_STATE, "set positions first");
    B2_TRY(dist_gather3(ctx, ctx->perdof[var]));
    return state_permute_to_user(ctx, ctx->perdof[var], values_dev);
}

extern "C" int b2_run(b2_context* ctx, int nsteps) {
    if (!ctx) return B2_ERR_ARG;
    return program_run(ctx, nsteps);
}

void phase_mark(b2_context* ctx, int tag) {
    if (!ctx->profiling) return;
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, ctx->stream);
    ctx->phase_events.push_back(e);
    ctx->phase_tags.push_back(tag);
}

// milliseconds per phase (B2_PHASE_* order, 6 values) accumulated since profiling was switched on
extern "C" int b2_get_phase_profile(b2_context* ctx, double out_ms[6]) {
    if (!ctx || !out_ms) return B2_ERR_ARG;
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < B2_PHASE_COUNT; k++) out_ms[k] = 0;
    for (size_t k = 0; k + 1 < ctx->phase_events.size(); k++) {
        float t = 0;
        B2_CUDA(cudaEventElapsedTime(&t, ctx->phase_events[k], ctx->phase_events[k+1]));
        const int tag = ctx->phase_tags[k];
        if (tag >= 0 && tag < B2_PHASE_COUNT) out_ms[tag] += t;
    }
    return B2_OK;
}

extern "C" int b2_set_profiling(b2_context* ctx, int on) {
    if (!ctx) return B2_ERR_ARG;
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    ctx->prof_events.clear();
    ctx->prof_tags.clear();
    for (cudaEvent_t e : ctx->phase_events) cudaEventDestroy(e);
    ctx->phase_events.clear();
    ctx->phase_tags.clear();
    ctx->profiling = on != 0;
    program_release(ctx);
    ctx->eager_steps = 0;
    return B2_OK;
}

extern "C" int b2_get_profile(b2_context* ctx, int handle, double* total_ms, long long* launches, long long* entries) {
    if (!ctx || handle < 0 || handle >= (int)ctx->pair_forces.size()) return b2_fail(ctx, B2_ERR_ARG, "bad pair force handle");
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    double ms = 0;
    long long count = 0;
    for (size_t k = 0; k < ctx->prof_tags.size(); k++) {
        if (ctx->prof_tags[k] != handle) continue;
        float t = 0;
        B2_CUDA(cudaEventElapsedTime(&t, ctx->prof_events[2*k], ctx->prof_events[2*k+1]));
        ms += t; count++;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = count;
    if (entries) {
        const NList& L = ctx->lists[ctx->pair_forces[handle].list];
        std::vector<int> c(ctx->ngroups);
        B2_CUDA(cudaMemcpy(c.data(), L.counts, sizeof(int)*ctx->ngroups, cudaMemcpyDeviceToHost));
        long long total = 0;
        for (int v : c) total += v;
        *entries = total;
    }
    return B2_OK;
}

extern "C" int b2_get_list_stats(b2_context* ctx, long long out_host[4]) {
    if (!ctx || !ctx->nl_flags) return B2_ERR_ARG;
    int flags[16];
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    B2_CUDA(cudaMemcpy(flags, ctx->nl_flags, sizeof(flags), cudaMemcpyDeviceToHost));
    out_host[0] = flags[2];      // rebuilds
    out_host[1] = flags[3];      // largest list
    out_host[2] = flags[8];      // fat groups at the last rebuild
    out_host[3] = ctx->ngroups;
    return B2_OK;
}

extern "C" int b2_get_counters(b2_context* ctx, long long out_host[8]) {
    if (!ctx) return B2_ERR_ARG;
    for (int k = 0; k < 8; k++) out_host[k] = ctx->counters[k];
    out_host[3] = ctx->nlists ? ctx->lists[0].cap : 0;
    return B2_OK;
}

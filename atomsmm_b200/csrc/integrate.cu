// Execution of lowered CustomIntegrator programs (kernels K5-K7 of SURVEY 2.1).
//
// Replaces OpenMM's CustomIntegrator interpreter (reference: integrators.py:163 hands the program
// over; propagators.py emits it).  Differences in execution strategy, not in results:
//   * counted while-loops are unrolled by the front end, data-dependent control flow over
//     globals runs inside ONE single-thread kernel (no host round trips, unlike OpenMM which
//     evaluates every condition on the host);
//   * force groups are cached per group against a position version, so the f2 -> f1 -> f0 -> f1
//     -> f2 pattern of RESPA never re-evaluates an unchanged group;
//   * kick / drift / scale steps have dedicated kernels (B2_OP_KICK/DRIFT/SCALE); anything else
//     goes through the per-DOF stack VM;
//   * the whole step is captured once into a CUDA graph and replayed.
#include <math.h>

#include "ctx.h"
#include "vm.cuh"

#define FULL 0xffffffffu

static PerDofTable make_table(b2_context* ctx) {
    PerDofTable t;
    memset(&t, 0, sizeof(t));
    t.vars[0] = ctx->x;
    t.vars[1] = ctx->v;
    for (size_t k = 0; k < ctx->perdof.size() && k < B2_MAX_PERDOF; k++) t.vars[2+k] = ctx->perdof[k];
    for (int g = 0; g < B2_FSLOTS; g++) t.f[g] = ctx->fbuf[g];
    t.mass = ctx->massd;
    return t;
}

__global__ void k_step_begin(unsigned long long* rng_state) { rng_state[2] += 1ull; }

__global__ void k_perdof(int dof_lo, int dof_hi, PerDofTable tab, int target, const int* __restrict__ code, int len,
                         const double* __restrict__ consts, double* globals,
                         const unsigned long long* __restrict__ rng_state, int serial, int mark_x) {
    const int dof = dof_lo + blockIdx.x*blockDim.x + threadIdx.x;
    if (dof >= dof_hi) return;
    RngStream rng;
    rng.seed = rng_state[0];
    rng.c0 = (uint32_t)dof; rng.c1 = (uint32_t)rng_state[2]; rng.c2 = (uint32_t)serial; rng.draw = 0;
    if (tab.mass[dof/3] == 0.0 && (target == 0 || target == 1)) return;   // massless particles are not moved
    const double value = vm_run<0>(code, len, consts, globals, &tab, dof, nullptr, &rng, nullptr);
    tab.vars[target][dof] = value;
}

__global__ void k_sum_partial(int dof_lo, int dof_hi, PerDofTable tab, const int* __restrict__ code, int len,
                              const double* __restrict__ consts, double* globals, double* partial) {
    __shared__ double sh[8];
    double s = 0;
    for (int dof = dof_lo + blockIdx.x*blockDim.x + threadIdx.x; dof < dof_hi; dof += gridDim.x*blockDim.x)
        s += vm_run<0>(code, len, consts, globals, &tab, dof, nullptr, nullptr, nullptr);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < (blockDim.x >> 5); k++) t += sh[k];
        partial[blockIdx.x] = t;
    }
}

// fast path for the ubiquitous  mvv <- sum(m*v*v)
__global__ void k_mvv_partial(int lo, int hi, const double* __restrict__ v, const double* __restrict__ mass, double* partial) {
    __shared__ double sh[8];
    double s = 0;
    for (int i = lo + blockIdx.x*blockDim.x + threadIdx.x; i < hi; i += gridDim.x*blockDim.x) {
        const double a = v[3*i], b = v[3*i+1], c = v[3*i+2];
        s += mass[i]*(a*a + b*b + c*c);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < (blockDim.x >> 5); k++) t += sh[k];
        partial[blockIdx.x] = t;
    }
}

__global__ void k_sum_final(int nblocks, const double* __restrict__ partial, double* globals, int target) {
    __shared__ double sh[256];
    double s = 0;
    for (int k = threadIdx.x; k < nblocks; k += blockDim.x) s += partial[k];   // fixed order: deterministic
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x/2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) globals[target] = sh[0];
}

// scalar program: one warp stages bytecode + constants in shared memory, lane 0 interprets
#define B2_GLOBAL_SMEM_INTS 2048
#define B2_GLOBAL_SMEM_CONSTS 256
__global__ void k_global(const int* __restrict__ code, int len, const double* __restrict__ consts, int nconsts,
                         double* globals, unsigned long long* rng_state, const double* energies) {
    __shared__ int scode[B2_GLOBAL_SMEM_INTS];
    __shared__ double sconsts[B2_GLOBAL_SMEM_CONSTS];
    const bool staged = 2*len <= B2_GLOBAL_SMEM_INTS && nconsts <= B2_GLOBAL_SMEM_CONSTS;
    if (staged) {
        for (int k = threadIdx.x; k < 2*len; k += blockDim.x) scode[k] = code[k];
        for (int k = threadIdx.x; k < nconsts; k += blockDim.x) sconsts[k] = consts[k];
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    RngStream rng;
    rng.seed = rng_state[0] ^ 0x5851f42d4c957f2dull;
    rng.c0 = 0xffffffffu; rng.c1 = (uint32_t)rng_state[1]; rng.c2 = (uint32_t)(rng_state[1] >> 32); rng.draw = 0;
    vm_run<1>(staged ? scode : code, len, staged ? sconsts : consts, globals, nullptr, 0, nullptr, &rng, energies);
    rng_state[1] += 1ull;
}

struct KickArgs {
    int nterms;
    const float4* f[B2_MAX_KICK_TERMS];
    int coef[B2_MAX_KICK_TERMS];
    float sign[B2_MAX_KICK_TERMS];
    int drift;                       // global index of the drift coefficient, -1: none
};

// v += sum_k s_k c_k f_k / m  [ ; x += c_d v ]      one thread per atom
__global__ void k_kick(int lo, int hi, double* __restrict__ v, double* __restrict__ x, KickArgs a,
                       const float* __restrict__ invm, const double* __restrict__ globals) {
    const int i = lo + blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const double im = (double)invm[i];
    if (im == 0.0) return;
    double ax = 0, ay = 0, az = 0;
#pragma unroll
    for (int k = 0; k < B2_MAX_KICK_TERMS; k++) {
        if (k >= a.nterms) break;
        const double c = globals[a.coef[k]]*(double)a.sign[k];
        const float4 f = a.f[k][i];
        ax += c*f.x; ay += c*f.y; az += c*f.z;
    }
    double vx = v[3*i] + im*ax, vy = v[3*i+1] + im*ay, vz = v[3*i+2] + im*az;
    v[3*i] = vx; v[3*i+1] = vy; v[3*i+2] = vz;
    if (a.drift >= 0) {
        const double c = globals[a.drift];
        x[3*i] += c*vx; x[3*i+1] += c*vy; x[3*i+2] += c*vz;
    }
}

__global__ void k_drift(int dof_lo, int dof_hi, double* __restrict__ x, const double* __restrict__ v,
                        const double* __restrict__ mass, const double* __restrict__ globals, int gcoef) {
    const int d = dof_lo + blockIdx.x*blockDim.x + threadIdx.x;
    if (d >= dof_hi) return;
    if (mass[d/3] == 0.0) return;
    x[d] += globals[gcoef]*v[d];
}

__global__ void k_scale(int dof_lo, int dof_hi, double* __restrict__ v, const double* __restrict__ globals, int gcoef) {
    const int d = dof_lo + blockIdx.x*blockDim.x + threadIdx.x;
    if (d >= dof_hi) return;
    v[d] *= globals[gcoef];
}

// ---------------------------------------------------------------------------------------------
// force-slot management
// ---------------------------------------------------------------------------------------------
int forces_ensure(b2_context* ctx, uint32_t mask, int slot) {
    if (slot < 0 || slot >= B2_FSLOTS) return b2_fail(ctx, B2_ERR_ARG, "bad force slot %d", slot);
    if (ctx->fbuf[slot] == nullptr) {
        B2_CUDA(cudaMalloc(&ctx->fbuf[slot], sizeof(float4)*ctx->n));
        ctx->fvalid[slot] = -1;
        ctx->graph_ready = false;
    }
    if (ctx->fvalid[slot] == ctx->pos_version) return B2_OK;
    bool any_pair = false;
    for (const PairForce& pf : ctx->pair_forces)
        if (mask & (1u << pf.group)) any_pair = true;
    if (any_pair) {
        B2_TRY(dist_sync_positions(ctx));
        B2_TRY(nl_prepare(ctx, false));
    }
    bool written = false;
    for (const PairForce& pf : ctx->pair_forces) {
        if (!(mask & (1u << pf.group))) continue;
        B2_TRY(pair_eval_forces(ctx, pf, ctx->fbuf[slot], written));
        written = true;
    }
    if (!written) B2_CUDA(cudaMemsetAsync(ctx->fbuf[slot], 0, sizeof(float4)*ctx->n, ctx->stream));
    B2_TRY(bonded_eval_forces(ctx, mask, ctx->fbuf[slot]));
    for (PmeForce& pm : ctx->pme_forces)
        if (mask & (1u << pm.group)) B2_TRY(pme_eval(ctx, pm, ctx->fbuf[slot], nullptr));
    ctx->fvalid[slot] = ctx->pos_version;
    return B2_OK;
}

// ---------------------------------------------------------------------------------------------
// program execution
// ---------------------------------------------------------------------------------------------
static int run_one_step(b2_context* ctx) {
    const int T = 256;
    const int lo = ctx->a_lo, hi = ctx->a_hi, n = hi - lo, ndof = 3*n;   // owned range
    cudaStream_t s = ctx->stream;
    k_step_begin<<<1, 1, 0, s>>>(ctx->rng_state);
    B2_LAUNCH_CHECK();
    for (const b2_op& op : ctx->ops) {
        switch (op.kind) {
        case B2_OP_EVAL:
            B2_TRY(forces_ensure(ctx, (uint32_t)op.a, op.b));
            break;
        case B2_OP_PERDOF: {
            PerDofTable tab = make_table(ctx);
            k_perdof<<<(ndof + T - 1)/T, T, 0, s>>>(3*lo, 3*hi, tab, op.a, ctx->code + op.b, op.c, ctx->consts,
                                                      ctx->globals, ctx->rng_state, op.e, 0);
            B2_LAUNCH_CHECK();
            if (op.a == 0) ctx->pos_version++;
            break;
        }
        case B2_OP_SUM: {
            const int blocks = 296;
            if (op.d == 1) {
                k_mvv_partial<<<blocks, T, 0, s>>>(lo, hi, ctx->v, ctx->massd, ctx->sum_partial);
            } else {
                PerDofTable tab = make_table(ctx);
                k_sum_partial<<<blocks, T, 0, s>>>(3*lo, 3*hi, tab, ctx->code + op.b, op.c, ctx->consts, ctx->globals,
                                                    ctx->sum_partial);
            }
            B2_LAUNCH_CHECK();
            k_sum_final<<<1, 256, 0, s>>>(blocks, ctx->sum_partial, ctx->globals, op.a);
            B2_LAUNCH_CHECK();
            B2_TRY(dist_allreduce(ctx, ctx->globals + op.a, 1));
            break;
        }
        case B2_OP_GLOBAL:
            k_global<<<1, 64, 0, s>>>(ctx->code + op.b, op.c, ctx->consts, ctx->nconsts, ctx->globals,
                                       ctx->rng_state, ctx->d_energy);
            B2_LAUNCH_CHECK();
            break;
        case B2_OP_KICK: {
            KickArgs ka;
            ka.nterms = op.a;
            ka.drift = op.c;
            if (op.a < 1 || op.a > B2_MAX_KICK_TERMS || op.b < 0 || op.b + 3*op.a > (int)ctx->h_code.size())
                return b2_fail(ctx, B2_ERR_ARG, "malformed kick op");
            for (int k = 0; k < B2_MAX_KICK_TERMS; k++) {
                const bool live = k < op.a;
                const int slot = live ? ctx->h_code[op.b + 3*k] : 0;
                if (live && (slot < 0 || slot >= B2_FSLOTS || ctx->fbuf[slot] == nullptr))
                    return b2_fail(ctx, B2_ERR_STATE, "kick uses force slot %d before it was evaluated", slot);
                ka.f[k] = live ? ctx->fbuf[slot] : nullptr;
                ka.coef[k] = live ? ctx->h_code[op.b + 3*k + 1] : 0;
                ka.sign[k] = live ? (float)ctx->h_code[op.b + 3*k + 2] : 0.f;
            }
            k_kick<<<(n + T - 1)/T, T, 0, s>>>(lo, hi, ctx->v, ctx->x, ka, ctx->invm, ctx->globals);
            B2_LAUNCH_CHECK();
            if (op.c >= 0) ctx->pos_version++;
            break;
        }
        case B2_OP_DRIFT:
            k_drift<<<(ndof + T - 1)/T, T, 0, s>>>(3*lo, 3*hi, ctx->x, ctx->v, ctx->massd, ctx->globals, op.a);
            B2_LAUNCH_CHECK();
            ctx->pos_version++;
            break;
        case B2_OP_SCALE:
            k_scale<<<(ndof + T - 1)/T, T, 0, s>>>(3*lo, 3*hi, ctx->v, ctx->globals, op.a);
            B2_LAUNCH_CHECK();
            break;
        case B2_OP_UPDATE_STATE:
            break;
        default:
            return b2_fail(ctx, B2_ERR_UNSUPPORTED, "unknown program op %d", op.kind);
        }
    }
    return B2_OK;
}

int program_release(b2_context* ctx) {
    if (ctx->graph_exec) {
        cudaGraphExecDestroy(ctx->graph_exec);
        ctx->graph_exec = nullptr;
    }
    ctx->graph_ready = false;
    return B2_OK;
}

static unsigned long long valid_mask(const b2_context* ctx) {
    unsigned long long m = 0;
    for (int g = 0; g < B2_FSLOTS; g++)
        if (ctx->fbuf[g] && ctx->fvalid[g] == ctx->pos_version) m |= 1ull << g;
    return m;
}

int program_run(b2_context* ctx, int nsteps) {
    if (!ctx->program_loaded) return b2_fail(ctx, B2_ERR_STATE, "no integrator program loaded");
    if (!ctx->have_positions) return b2_fail(ctx, B2_ERR_STATE, "positions have not been set");
    // Which force slots are valid is tracked on the host against pos_version.  A captured graph
    // replays the launch sequence of a steady-state step, so it may only be launched when every
    // slot it assumed valid at entry is valid now; otherwise one step runs eagerly, which
    // restores the steady-state pattern.  The first step always runs eagerly (it also performs
    // all lazy allocations, which are illegal during capture).
    static const bool graph_allowed = getenv("B2_NO_GRAPH") == nullptr;
    const bool use_graph = graph_allowed && !ctx->profiling;
    for (int done = 0; done < nsteps; done++) {
        const unsigned long long entry = valid_mask(ctx);
        const bool synced = ctx->x_synced == ctx->pos_version;
        if (use_graph && ctx->graph_ready && (entry & ctx->graph_entry_mask) == ctx->graph_entry_mask &&
            (synced || !ctx->graph_entry_synced)) {
            B2_CUDA(cudaGraphLaunch(ctx->graph_exec, ctx->stream));
            ctx->counters[5]++;
            ctx->pos_version += ctx->graph_dpos;
            if (ctx->graph_exit_synced) ctx->x_synced = ctx->pos_version;
            for (int g = 0; g < B2_FSLOTS; g++)
                if (ctx->graph_exit_mask & (1ull << g)) ctx->fvalid[g] = ctx->pos_version;
            continue;
        }
        if (use_graph && !ctx->graph_ready && ctx->eager_steps >= 1) {
            const long long v0 = ctx->pos_version;
            const long long launches0 = ctx->counters[0];
            B2_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            const int r = run_one_step(ctx);
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
            if (r != B2_OK) { if (graph) cudaGraphDestroy(graph); return r; }
            if (e != cudaSuccess) return b2_fail(ctx, B2_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
            e = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return b2_fail(ctx, B2_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(e));
            ctx->graph_entry_mask = entry;
            ctx->graph_entry_synced = synced;
            ctx->graph_exit_synced = ctx->x_synced == ctx->pos_version;
            ctx->graph_exit_mask = valid_mask(ctx);
            ctx->graph_dpos = ctx->pos_version - v0;
            ctx->counters[6] = ctx->counters[0] - launches0;   // kernels per step
            ctx->graph_ready = true;
            B2_CUDA(cudaGraphLaunch(ctx->graph_exec, ctx->stream));
            ctx->counters[5]++;
            continue;
        }
        B2_TRY(run_one_step(ctx));
        ctx->eager_steps++;
    }
    return B2_OK;
}

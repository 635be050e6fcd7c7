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

#include <stdio.h>

#include <algorithm>
#include <chrono>

#include "bonded.cuh"
#include "dd.cuh"

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

// after v <- s v the carried sum(m v.v) is s^2 times its old value
__global__ void k_mvv_rescale(double* globals, int mvv, int factor) { globals[mvv] *= globals[factor]*globals[factor]; }

__global__ void k_fold_derivatives(double* e) { e[64] += e[74]; e[65] += e[75]; }

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

struct KickArgs {
    int nterms;
    const float4* f[B2_MAX_KICK_TERMS];
    int coef[B2_MAX_KICK_TERMS];
    float sign[B2_MAX_KICK_TERMS];
    int drift;                       // global index of the drift coefficient, -1: none
    int prescale;                    // global index of a factor applied to v first, -1: none
    int mvv;                         // global index receiving sum(m*v*v) of the new velocities, -1: none
    int code_len;                    // scalar program run once after the sum (0: none)
};

// Scalar program executed by one thread, with bytecode, constants AND the global variables staged in
// shared memory by the whole block (every VM instruction is a dependent load: from HBM/L2 that is
// ~0.3 us per instruction, from shared memory ~30 ns).  Falls back to global memory when it does not fit.
#define B2_STAGE_INTS 1024
#define B2_STAGE_CONSTS 256
#define B2_STAGE_GLOBALS 256
struct ScalarStage {
    int code[B2_STAGE_INTS];
    double consts[B2_STAGE_CONSTS];
    double globals[B2_STAGE_GLOBALS];
};

__device__ void run_scalar_program_staged(ScalarStage& st, const int* __restrict__ code, int len,
                                          const double* __restrict__ consts, int nconsts, double* globals,
                                          int nglobals, unsigned long long* rng_state, const double* energies) {
    const bool staged = 2*len <= B2_STAGE_INTS && nconsts <= B2_STAGE_CONSTS && nglobals <= B2_STAGE_GLOBALS;
    if (staged) {
        for (int k = threadIdx.x; k < 2*len; k += blockDim.x) st.code[k] = code[k];
        for (int k = threadIdx.x; k < nconsts; k += blockDim.x) st.consts[k] = consts[k];
        for (int k = threadIdx.x; k < nglobals; k += blockDim.x) st.globals[k] = __ldcg(&globals[k]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        RngStream rng;
        rng.seed = rng_state[0] ^ 0x5851f42d4c957f2dull;
        rng.c0 = 0xffffffffu; rng.c1 = (uint32_t)rng_state[1]; rng.c2 = (uint32_t)(rng_state[1] >> 32); rng.draw = 0;
        if (staged) vm_run<1>(st.code, len, st.consts, st.globals, nullptr, 0, nullptr, &rng, energies);
        else vm_run<1>(code, len, consts, globals, nullptr, 0, nullptr, &rng, energies);
        rng_state[1] += 1ull;
    }
    __syncthreads();
    if (staged)
        for (int k = threadIdx.x; k < nglobals; k += blockDim.x) globals[k] = st.globals[k];
}

// stand-alone scalar program (prologue of a step, Bussi's rejection loop, AFED wall reflection ...)
__global__ void k_global(const int* __restrict__ code, int len, const double* __restrict__ consts, int nconsts,
                         double* globals, int nglobals, unsigned long long* rng_state, const double* energies,
                         int begin_step) {
    __shared__ ScalarStage stage;
    if (begin_step && threadIdx.x == 0) rng_state[2] += 1ull;      // MD step counter of the RNG streams
    run_scalar_program_staged(stage, code, len, consts, nconsts, globals, nglobals, rng_state, energies);
}

// Domain decomposition, peer-memory mode: the all-reduce of a thermostat's sum(m v.v) and the scalar
// program that consumes it are ONE single-block kernel -- partials are pushed into the peers' signal
// blocks over NVLink and added in rank order (dd.cuh), no library collective in the step graph.
__global__ void k_global_reduce(DDPeers P, unsigned long long* dd_state, int target, const int* __restrict__ code, int len,
                                const double* __restrict__ consts, int nconsts, double* globals, int nglobals,
                                unsigned long long* rng_state, const double* energies) {
    __shared__ ScalarStage stage;
    __shared__ double part[B2_MAX_RANKS];
    __shared__ unsigned long long epoch;
    const double total = dd_allreduce_sum(P, dd_state, __ldcg(&globals[target]), threadIdx.x, part, &epoch);
    __syncthreads();
    if (threadIdx.x == 0) { globals[target] = total; __threadfence(); }
    __syncthreads();
    if (len > 0) run_scalar_program_staged(stage, code, len, consts, nconsts, globals, nglobals, rng_state, energies);
}

// The velocity kernel: v <- s*v + sum_k s_k c_k f_k/m ; [x += c_d v] ; [mvv <- sum(m v.v) ; scalar
// program].  One thread per owned atom.  The reduction is deterministic: fixed tree inside a block,
// per-block partials summed in index order by the last block to finish, which then also runs the
// scalar program that consumes the sum (a Nose-Hoover update), so a thermostat block is ONE launch.
template <bool MVV>
__global__ void __launch_bounds__(256) k_vel(int lo, int hi, double* __restrict__ v, double* __restrict__ x, KickArgs a,
                                             const double* __restrict__ mass, double* globals,
                                             double* __restrict__ partial, unsigned* __restrict__ ticket,
                                             const int* __restrict__ code, const double* __restrict__ consts,
                                             int nconsts, int nglobals,
                                             unsigned long long* rng_state, const double* energies) {
    const int i = lo + blockIdx.x*blockDim.x + threadIdx.x;
    double mvv = 0;
    if (i < hi) {
        const double m = mass[i];
        if (m != 0.0) {
            double vx = v[3*i], vy = v[3*i+1], vz = v[3*i+2];
            if (a.prescale >= 0) {
                const double s = globals[a.prescale];
                vx *= s; vy *= s; vz *= s;
            }
            if (a.nterms > 0) {
                double ax = 0, ay = 0, az = 0;
#pragma unroll
                for (int k = 0; k < B2_MAX_KICK_TERMS; k++) {
                    if (k < a.nterms) {
                        const float4 f = a.f[k][i];
                        const double c = (double)a.sign[k]*globals[a.coef[k]];
                        ax += c*(double)f.x; ay += c*(double)f.y; az += c*(double)f.z;
                    }
                }
                const double im = 1.0/m;
                vx += ax*im; vy += ay*im; vz += az*im;
            }
            if (a.prescale >= 0 || a.nterms > 0) { v[3*i] = vx; v[3*i+1] = vy; v[3*i+2] = vz; }
            if (a.drift >= 0) {
                const double c = globals[a.drift];
                x[3*i] += c*vx; x[3*i+1] += c*vy; x[3*i+2] += c*vz;
            }
            mvv = m*(vx*vx + vy*vy + vz*vz);
        }
    }
    if constexpr (MVV) {
    __shared__ double sh[8];
    __shared__ bool last;
    for (int o = 16; o > 0; o >>= 1) mvv += __shfl_xor_sync(FULL, mvv, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mvv;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < (blockDim.x >> 5); k++) t += sh[k];
        partial[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    __shared__ double red[256];
    double s = 0;
    for (int k = threadIdx.x; k < gridDim.x; k += blockDim.x) s += __ldcg(&partial[k]);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x/2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        globals[a.mvv] = red[0];
        *ticket = 0;
        __threadfence();
    }
    __syncthreads();
    if (a.code_len > 0) {
        __shared__ ScalarStage stage;
        run_scalar_program_staged(stage, code, a.code_len, consts, nconsts, globals, nglobals, rng_state, energies);
    }
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
// Fused RESPA inner loop (north-star kernel K5): a run of  kick[+drift] / evaluate-bonded-forces
// ops is ONE launch.  All explicit-list (bonded) terms are intramolecular and molecules are
// contiguous in the engine's order, so a thread block that owns a chunk of whole molecules can
// iterate  v += c f/m ; x += c v ; f0 <- bonded(x)  entirely on chip: positions and the inner
// force live in shared memory, x / v / outer forces are read once and x / v / f0 written once per
// run instead of once per inner iteration.
// ---------------------------------------------------------------------------------------------
#define B2_INNER_MAX_OPS 40
struct InnerOp {
    int kind;                        // 0 kick, 1 evaluate the local (bonded) force
    int nterms;
    int slot[B2_MAX_KICK_TERMS];
    int coef[B2_MAX_KICK_TERMS];
    float sign[B2_MAX_KICK_TERMS];
    int drift, prescale;
};

struct InnerArgs {
    int nops;
    int local_slot;                  // force slot produced by the local evaluations
    int write_force;                 // the last local evaluation is valid for the final positions
    InnerOp op[B2_INNER_MAX_OPS];
    const float4* f[B2_FSLOTS];
    int nforces;                     // bonded forces taking part
    BondArgs a[B2_MAX_BATCH];
    int arity[B2_MAX_BATCH];
    int batch_of[B2_INNER_MAX_FORCES];   // bonded-force index -> entry of a[] or -1
};

// tuning knobs (measured on B200: profiles/round2_pair_variants.txt)
#ifndef B2_INNER_MINB
#define B2_INNER_MINB 8         // resident blocks per SM the register budget is sized for (80 registers at 96 threads)
#endif
#ifndef B2_INNER_HOIST
#define B2_INNER_HOIST 1        // resolve each thread's first term once per launch
#endif
template <bool CUSTOM>
__global__ void __launch_bounds__(B2_CHUNK, B2_INNER_MINB) k_inner(const int* __restrict__ chunk_start,
                                                    const int* __restrict__ term_ptr,
                                                    const int2* __restrict__ terms, double* __restrict__ xg,
                                                    double* __restrict__ vg, const double* __restrict__ mass,
                                                    const double* __restrict__ globals, float4* __restrict__ fout,
                                                    const __grid_constant__ InnerArgs A) {
    __shared__ double xs[3*B2_CHUNK];
    __shared__ unsigned long long fs[3*B2_CHUNK];
    const int c = blockIdx.x, tid = threadIdx.x;
    const int base = chunk_start[c], count = chunk_start[c+1] - base;
    const int i = base + tid;
    const bool mine = tid < count;
    double x[3] = {0, 0, 0}, v[3] = {0, 0, 0}, im = 0;
    if (mine) {
        const double m = mass[i];
        im = m != 0.0 ? 1.0/m : 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) { x[k] = xg[3*i+k]; v[k] = vg[3*i+k]; }
    }
    const int t0 = term_ptr[c], t1 = term_ptr[c+1];
    const LocalGeo geo{xs, fs, base};
    bool have_local = false;
    // This thread's first term, resolved once for the whole run when it belongs to a closed-form family (harmonic
    // bond / angle, periodic torsion: <= 3 parameters, which travel in registers too).  The chunk's terms are dealt out
    // one per thread; further terms of a thread (large molecules) and other families are resolved at every
    // evaluation, as before.
    int h_b = -1, h_ar = 0, h_a0 = 0, h_a1 = 0, h_a2 = 0, h_a3 = 0;
    double h_p[3] = {0, 0, 0};
    if (B2_INNER_HOIST && t0 + tid < t1) {
        const int2 rec = terms[t0 + tid];
        const int b = A.batch_of[rec.x];
        if (b >= 0) {
            const BondArgs& ba = A.a[b];
            if (ba.family == B2_BOND_HARMONIC || ba.family == B2_ANGLE_HARMONIC || ba.family == B2_TORSION_PERIODIC) {
                h_ar = A.arity[b];
                const int* at = ba.atoms + (size_t)h_ar*rec.y;
                h_a0 = ba.inv[at[0]];
                h_a1 = ba.inv[at[1]];
                if (h_ar > 2) h_a2 = ba.inv[at[2]];
                if (h_ar > 3) h_a3 = ba.inv[at[3]];
                const double* p = ba.params + (size_t)rec.y*ba.stride;
                h_p[0] = p[0]; h_p[1] = p[1];
                if (ba.stride > 2) h_p[2] = p[2];
                h_b = (h_a0 >= ba.a_lo && h_a0 < ba.a_hi) ? b : -2;      // -2: resolved, owned by another rank
            }
        }
    }
    const int t_first = h_b == -1 ? t0 + tid : t0 + tid + B2_CHUNK;     // where the per-evaluation loop starts
    for (int q = 0; q < A.nops; q++) {
        const InnerOp& op = A.op[q];
        if (op.kind == 0) {
            if (mine && im != 0.0) {
                if (op.prescale >= 0) {
                    const double s = globals[op.prescale];
                    v[0] *= s; v[1] *= s; v[2] *= s;
                }
                double a[3] = {0, 0, 0};
                for (int k = 0; k < op.nterms; k++) {
                    const double cf = (double)op.sign[k]*globals[op.coef[k]];
                    if (have_local && op.slot[k] == A.local_slot) {
                        a[0] += cf*geo.force(tid, 0); a[1] += cf*geo.force(tid, 1); a[2] += cf*geo.force(tid, 2);
                    } else {
                        const float4 f = A.f[op.slot[k]][i];
                        a[0] += cf*(double)f.x; a[1] += cf*(double)f.y; a[2] += cf*(double)f.z;
                    }
                }
#pragma unroll
                for (int k = 0; k < 3; k++) v[k] += a[k]*im;
                if (op.drift >= 0) {
                    const double cd = globals[op.drift];
#pragma unroll
                    for (int k = 0; k < 3; k++) x[k] += cd*v[k];
                }
            }
        } else {
            __syncthreads();           // everybody has consumed the previous local force
#pragma unroll
            for (int k = 0; k < 3; k++) { xs[3*tid+k] = x[k]; fs[3*tid+k] = 0ull; }
            __syncthreads();
            if (h_b >= 0) {
                const BondArgs& ba = A.a[h_b];
                double e = 0, w = 0;
                if (h_ar == 2) term_bond2_core<true, false, false>(ba, h_a0, h_a1, h_p, geo, e, w);
                else if (h_ar == 3) term_angle_core<true, false, false>(ba, h_a0, h_a1, h_a2, h_p, geo, e);
                else term_torsion_core<true, false>(ba, h_a0, h_a1, h_a2, h_a3, h_p, geo, e);
            }
            for (int t = t_first; t < t1; t += B2_CHUNK) {
                const int2 rec = terms[t];
                const int b = A.batch_of[rec.x];
                if (b < 0) continue;
                double e = 0, w = 0;
                if (A.arity[b] == 2) term_bond2<true, false, CUSTOM>(A.a[b], rec.y, geo, e, w);
                else if (A.arity[b] == 3) term_angle<true, false, CUSTOM>(A.a[b], rec.y, geo, e);
                else term_torsion<true, false>(A.a[b], rec.y, geo, e);
            }
            __syncthreads();
            have_local = true;
        }
    }
    if (mine) {
#pragma unroll
        for (int k = 0; k < 3; k++) { xg[3*i+k] = x[k]; vg[3*i+k] = v[k]; }
        if (A.write_force) fout[i] = make_float4((float)geo.force(tid, 0), (float)geo.force(tid, 1), (float)geo.force(tid, 2), 0.f);
    }
}

// Chunks of whole molecules (<= B2_CHUNK atoms) over the owned range, and for every chunk the
// bonded terms of its molecules as (bonded-force index, term index) records.
int inner_prepare(b2_context* ctx) {
    if (ctx->inner_built) return B2_OK;
    ctx->inner_built = true;
    ctx->inner_ok = false;
    cudaFree(ctx->chunk_start); cudaFree(ctx->chunk_term_ptr); cudaFree(ctx->chunk_terms);
    ctx->chunk_start = nullptr; ctx->chunk_term_ptr = nullptr; ctx->chunk_terms = nullptr;
    if ((int)ctx->bonded_forces.size() > B2_INNER_MAX_FORCES || ctx->h_orig.empty()) return B2_OK;
    const int lo = ctx->a_lo, hi = ctx->a_hi;
    std::vector<int> start;
    int s = lo;
    while (s < hi) {
        start.push_back(s);
        int e = s;
        while (e < hi) {
            int m_end = e + 1;      // end of the molecule starting at e
            while (m_end < hi && ctx->h_mol[ctx->h_orig[m_end]] == ctx->h_mol[ctx->h_orig[e]]) m_end++;
            if (m_end - s > B2_CHUNK) break;
            e = m_end;
        }
        if (e == s) return B2_OK;   // a molecule larger than a chunk: no fused path
        s = e;
    }
    start.push_back(hi);
    const int nchunks = (int)start.size() - 1;
    if (nchunks <= 0) return B2_OK;
    std::vector<int> inv(ctx->n), chunk_of(ctx->n, -1);
    for (int k = 0; k < ctx->n; k++) inv[ctx->h_orig[k]] = k;
    for (int c = 0; c < nchunks; c++)
        for (int k = start[c]; k < start[c+1]; k++) chunk_of[k] = c;
    std::vector<int> ptr(nchunks + 1, 0);
    for (size_t f = 0; f < ctx->bonded_forces.size(); f++) {
        const BondedForce& bf = ctx->bonded_forces[f];
        for (int t = 0; t < bf.nterms; t++) {
            const int c = chunk_of[inv[bf.h_atoms[(size_t)bf.arity*t]]];
            if (c < 0) continue;     // owned by another rank
            for (int q = 1; q < bf.arity; q++)
                if (chunk_of[inv[bf.h_atoms[(size_t)bf.arity*t + q]]] != c) return B2_OK;   // not intramolecular
            ptr[c+1]++;
        }
    }
    for (int c = 0; c < nchunks; c++) ptr[c+1] += ptr[c];
    std::vector<int2> recs(std::max(1, ptr[nchunks]));
    std::vector<int> cur(ptr.begin(), ptr.end() - 1);
    for (size_t f = 0; f < ctx->bonded_forces.size(); f++) {
        const BondedForce& bf = ctx->bonded_forces[f];
        for (int t = 0; t < bf.nterms; t++) {
            const int c = chunk_of[inv[bf.h_atoms[(size_t)bf.arity*t]]];
            if (c < 0) continue;
            recs[cur[c]++] = make_int2((int)f, t);
        }
    }
    B2_CUDA(cudaMalloc(&ctx->chunk_start, sizeof(int)*(nchunks + 1)));
    B2_CUDA(cudaMalloc(&ctx->chunk_term_ptr, sizeof(int)*(nchunks + 1)));
    B2_CUDA(cudaMalloc(&ctx->chunk_terms, sizeof(int2)*recs.size()));
    B2_CUDA(cudaMemcpy(ctx->chunk_start, start.data(), sizeof(int)*(nchunks + 1), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(ctx->chunk_term_ptr, ptr.data(), sizeof(int)*(nchunks + 1), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(ctx->chunk_terms, recs.data(), sizeof(int2)*recs.size(), cudaMemcpyHostToDevice));
    ctx->nchunks = nchunks;
    ctx->inner_ok = true;
    return B2_OK;
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
        if (!ctx->p2p) B2_TRY(dist_sync_positions(ctx));     // peer-memory mode: the halo arrives inside nl_prepare
        B2_TRY(nl_prepare(ctx, false));
    }
    bool written = false;
    phase_mark(ctx, B2_PHASE_PAIR);
    for (const PairForce& pf : ctx->pair_forces) {
        if (!(mask & (1u << pf.group))) continue;
        B2_TRY(pair_eval_forces(ctx, pf, ctx->fbuf[slot], written));
        written = true;
    }
    phase_mark(ctx, B2_PHASE_INTEGRATE);
    B2_TRY(dist_before_move(ctx));      // by now the peers have long finished reading: the wait costs nothing here
    if (!written) B2_CUDA(cudaMemsetAsync(ctx->fbuf[slot], 0, sizeof(float4)*ctx->n, ctx->stream));
    B2_TRY(bonded_eval_forces(ctx, mask, ctx->fbuf[slot]));
    for (PmeForce& pm : ctx->pme_forces)
        if (mask & (1u << pm.group)) B2_TRY(pme_eval(ctx, pm, ctx->fbuf[slot], nullptr));
    ctx->fvalid[slot] = ctx->pos_version;
    return B2_OK;
}

static bool has_pair_force(const b2_context* ctx, uint32_t mask) {
    for (const PairForce& pf : ctx->pair_forces)
        if (mask & (1u << pf.group)) return true;
    return false;
}

// Two force slots that both need pair work at the same positions (RESPA evaluates the near and the
// far force back to back at the end of a step): their pair kernels are independent -- different
// lists, different outputs -- and run concurrently on two streams, which become two branches of
// the step's CUDA graph.  Each kernel alone leaves ~30 % of the issue slots idle and has its own tail.
static int forces_ensure_dual(b2_context* ctx, uint32_t mask_a, int slot_a, uint32_t mask_b, int slot_b) {
    if (!ctx->p2p) B2_TRY(dist_sync_positions(ctx));
    B2_TRY(nl_prepare(ctx, false));
    B2_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    B2_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
    bool written_a = false, written_b = false;
    for (const PairForce& pf : ctx->pair_forces) {
        if (!(mask_b & (1u << pf.group))) continue;
        B2_TRY(pair_eval_forces(ctx, pf, ctx->fbuf[slot_b], written_b, 1));
        written_b = true;
    }
    for (const PairForce& pf : ctx->pair_forces) {
        if (!(mask_a & (1u << pf.group))) continue;
        B2_TRY(pair_eval_forces(ctx, pf, ctx->fbuf[slot_a], written_a, 0));
        written_a = true;
    }
    B2_CUDA(cudaEventRecord(ctx->ev_join, ctx->side_stream));
    B2_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    B2_TRY(dist_before_move(ctx));
    const uint32_t masks[2] = {mask_a, mask_b};
    const int slots[2] = {slot_a, slot_b};
    for (int k = 0; k < 2; k++) {
        B2_TRY(bonded_eval_forces(ctx, masks[k], ctx->fbuf[slots[k]]));
        for (PmeForce& pm : ctx->pme_forces)
            if (masks[k] & (1u << pm.group)) B2_TRY(pme_eval(ctx, pm, ctx->fbuf[slots[k]], nullptr));
        ctx->fvalid[slots[k]] = ctx->pos_version;
    }
    return B2_OK;
}

// ---------------------------------------------------------------------------------------------
// program execution
// ---------------------------------------------------------------------------------------------
static bool bonded_only(const b2_context* ctx, uint32_t mask) {
    for (const PairForce& pf : ctx->pair_forces)
        if (mask & (1u << pf.group)) return false;
    for (const PmeForce& pm : ctx->pme_forces)
        if (mask & (1u << pm.group)) return false;
    return true;
}

static int ensure_partials(b2_context* ctx, int blocks) {
    if (blocks <= ctx->sum_partial_size) return B2_OK;
    cudaFree(ctx->sum_partial);
    ctx->sum_partial = nullptr;
    B2_CUDA(cudaMalloc(&ctx->sum_partial, sizeof(double)*blocks));
    ctx->sum_partial_size = blocks;
    return B2_OK;
}

static int fill_kick(b2_context* ctx, const b2_op& op, KickArgs& ka) {
    memset(&ka, 0, sizeof(ka));
    ka.nterms = op.a;
    ka.drift = op.c;
    ka.prescale = op.d;
    ka.mvv = op.e;
    ka.code_len = op.g;
    if (op.a < 0 || op.a > B2_MAX_KICK_TERMS || op.b < 0 || op.b + 3*op.a > (int)ctx->h_code.size())
        return b2_fail(ctx, B2_ERR_ARG, "malformed kick op");
    for (int k = 0; k < op.a; k++) {
        const int slot = ctx->h_code[op.b + 3*k];
        if (slot < 0 || slot >= B2_FSLOTS || ctx->fbuf[slot] == nullptr)
            return b2_fail(ctx, B2_ERR_STATE, "kick uses force slot %d before it was evaluated", slot);
        ka.f[k] = ctx->fbuf[slot];
        ka.coef[k] = ctx->h_code[op.b + 3*k + 1];
        ka.sign[k] = (float)ctx->h_code[op.b + 3*k + 2];
    }
    return B2_OK;
}

static int launch_vel(b2_context* ctx, const b2_op& op) {
    const int T = 256, lo = ctx->a_lo, hi = ctx->a_hi;
    const int blocks = std::max(1, (hi - lo + T - 1)/T);
    KickArgs ka;
    B2_TRY(fill_kick(ctx, op, ka));
    if (op.c >= 0) B2_TRY(dist_before_move(ctx));
    cudaStream_t s = ctx->stream;
    static const bool carry_allowed = getenv("B2_NO_MVV_CARRY") == nullptr;
    const bool changes_v = ka.nterms > 0 || ka.prescale >= 0;
    // (a) a pure "sum(m v.v), then the scalar program" block whose sum is already known: only the program runs
    if (carry_allowed && ka.mvv >= 0 && !changes_v && ka.drift < 0 && ka.mvv == ctx->mvv_index &&
        ctx->mvv_version == ctx->v_version) {
        if (op.g > 0) {
            k_global<<<1, 64, 0, s>>>(ctx->code + op.f, op.g, ctx->consts, ctx->nconsts, ctx->globals,
                                       ctx->nglobals, ctx->rng_state, ctx->d_energy, 0);
            B2_LAUNCH_CHECK();
        }
        return B2_OK;
    }
    // with several ranks the sum must be all-reduced before the scalar program may consume it
    const bool split = ctx->nranks > 1 && ka.mvv >= 0;
    if (split) ka.code_len = 0;
    if (ka.mvv >= 0) {
        k_vel<true><<<blocks, T, 0, s>>>(lo, hi, ctx->v, ctx->x, ka, ctx->massd, ctx->globals, ctx->sum_partial,
                                          ctx->ticket, ctx->code + op.f, ctx->consts, ctx->nconsts, ctx->nglobals,
                                          ctx->rng_state, ctx->d_energy);
    } else {
        k_vel<false><<<blocks, T, 0, s>>>(lo, hi, ctx->v, ctx->x, ka, ctx->massd, ctx->globals, ctx->sum_partial,
                                           ctx->ticket, ctx->code + op.f, ctx->consts, ctx->nconsts, ctx->nglobals,
                                          ctx->rng_state, ctx->d_energy);
    }
    B2_LAUNCH_CHECK();
    if (split) phase_mark(ctx, B2_PHASE_REDUCE);
    if (split && ctx->p2p) {
        DDPeers P;
        dd_fill_peers(ctx, &P);
        k_global_reduce<<<1, 64, 0, s>>>(P, ctx->dd_state, ka.mvv, ctx->code + op.f, op.g, ctx->consts, ctx->nconsts,
                                         ctx->globals, ctx->nglobals, ctx->rng_state, ctx->d_energy);
        B2_LAUNCH_CHECK();
    } else if (split) {
        B2_TRY(dist_allreduce(ctx, ctx->globals + ka.mvv, 1));
        if (op.g > 0) {
            k_global<<<1, 64, 0, s>>>(ctx->code + op.f, op.g, ctx->consts, ctx->nconsts, ctx->globals,
                                       ctx->nglobals, ctx->rng_state, ctx->d_energy, 0);
            B2_LAUNCH_CHECK();
        }
    }
    if (split) phase_mark(ctx, B2_PHASE_INTEGRATE);
    if (op.c >= 0) ctx->pos_version++;
    // bookkeeping of the carried sum
    const bool was_valid = ctx->mvv_index >= 0 && ctx->mvv_version == ctx->v_version;
    if (changes_v) ctx->v_version++;
    if (ka.mvv >= 0) {                                   // the kernel summed the NEW velocities
        ctx->mvv_index = ka.mvv;
        ctx->mvv_version = ctx->v_version;
    } else if (carry_allowed && was_valid && ka.nterms == 0 && ka.prescale >= 0 && ctx->mvv_factor_hint >= 0) {
        // (b) pure rescaling v <- s v announced by B2_OP_MVV_FACTOR: the carried sum follows
        k_mvv_rescale<<<1, 1, 0, s>>>(ctx->globals, ctx->mvv_index, ctx->mvv_factor_hint);
        B2_LAUNCH_CHECK();
        ctx->mvv_version = ctx->v_version;
    }
    ctx->mvv_factor_hint = -1;
    return B2_OK;
}

// Try to execute ops[k...] as one fused inner-loop launch.  Returns the number of ops consumed
// (0: not applicable, run ops[k] the ordinary way).
static int try_fused_run(b2_context* ctx, size_t k, int* consumed) {
    *consumed = 0;
    static const bool allowed = getenv("B2_NO_FUSED_INNER") == nullptr;
    if (!allowed || !ctx->inner_ok) return B2_OK;
    const std::vector<b2_op>& ops = ctx->ops;
    if (ops[k].kind != B2_OP_KICK || ops[k].e >= 0) return B2_OK;
    // the run: KICK (no reduction) and bonded-only EVAL ops, all evaluations into the same slot
    InnerArgs A;
    A.nops = 0; A.local_slot = -1; A.write_force = 0;
    uint32_t local_mask = 0;
    long long version = ctx->pos_version;
    long long local_valid = -1;          // position version the local force belongs to
    int evals = 0;
    size_t q = k;
    for (; q < ops.size() && A.nops < B2_INNER_MAX_OPS; q++) {
        const b2_op& op = ops[q];
        if (op.kind == B2_OP_KICK) {
            if (op.e >= 0) break;        // carries a reduction: ordinary velocity kernel
            KickArgs ka;
            B2_TRY(fill_kick(ctx, op, ka));
            InnerOp& io = A.op[A.nops];
            io.kind = 0; io.nterms = ka.nterms; io.drift = ka.drift; io.prescale = ka.prescale;
            bool ok = true;
            for (int t = 0; t < ka.nterms; t++) {
                io.slot[t] = ctx->h_code[op.b + 3*t]; io.coef[t] = ka.coef[t]; io.sign[t] = ka.sign[t];
                // every term must be valid at this point: locally (the kernel then reads shared
                // memory) or in its global buffer
                const bool local = io.slot[t] == A.local_slot && local_valid == version;
                if (!local && (io.slot[t] == A.local_slot || ctx->fvalid[io.slot[t]] != version)) ok = false;
            }
            if (!ok) break;
            A.nops++;
            if (op.c >= 0) version++;
        } else if (op.kind == B2_OP_EVAL) {
            const uint32_t mask = (uint32_t)op.a;
            const bool valid = (op.b == A.local_slot && local_valid == version) ||
                               (op.b != A.local_slot && ctx->fbuf[op.b] && ctx->fvalid[op.b] == version);
            if (valid) continue;         // nothing to do; the op stays inside the run
            if (!bonded_only(ctx, mask)) break;
            if (A.local_slot >= 0 && (op.b != A.local_slot || mask != local_mask)) break;
            if (ctx->fbuf[op.b] == nullptr) break;    // first use allocates (ordinary path)
            A.local_slot = op.b; local_mask = mask;
            A.op[A.nops].kind = 1;
            A.nops++;
            local_valid = version;
            evals++;
        } else {
            break;
        }
    }
    if (evals == 0) return B2_OK;        // a lone kick: the ordinary velocity kernel
    // bonded forces of the local mask
    A.nforces = 0;
    for (int f = 0; f < B2_INNER_MAX_FORCES; f++) A.batch_of[f] = -1;
    for (size_t f = 0; f < ctx->bonded_forces.size(); f++) {
        const BondedForce& bf = ctx->bonded_forces[f];
        if (!(local_mask & (1u << bf.group)) || bf.nterms == 0) continue;
        if ((bf.family == B2_BOND_CUSTOM || bf.family == B2_ANGLE_CUSTOM) && bf.ncode_de == 0) continue;
        if (A.nforces == B2_MAX_BATCH) return B2_OK;
        A.a[A.nforces] = bonded_make_args(ctx, bf);
        A.arity[A.nforces] = bf.arity;
        A.batch_of[f] = A.nforces++;
    }
    for (int g = 0; g < B2_FSLOTS; g++) A.f[g] = ctx->fbuf[g];
    A.write_force = local_valid == version ? 1 : 0;
    B2_TRY(dist_before_move(ctx));
    bool custom = false;
    for (size_t f = 0; f < ctx->bonded_forces.size(); f++)
        if (A.batch_of[f] >= 0 && (ctx->bonded_forces[f].family == B2_BOND_CUSTOM || ctx->bonded_forces[f].family == B2_ANGLE_CUSTOM))
            custom = true;
    if (custom)
        k_inner<true><<<ctx->nchunks, B2_CHUNK, 0, ctx->stream>>>(ctx->chunk_start, ctx->chunk_term_ptr, ctx->chunk_terms,
                                                                  ctx->x, ctx->v, ctx->massd, ctx->globals,
                                                                  ctx->fbuf[A.local_slot], A);
    else
        k_inner<false><<<ctx->nchunks, B2_CHUNK, 0, ctx->stream>>>(ctx->chunk_start, ctx->chunk_term_ptr, ctx->chunk_terms,
                                                                   ctx->x, ctx->v, ctx->massd, ctx->globals,
                                                                   ctx->fbuf[A.local_slot], A);
    B2_LAUNCH_CHECK();
    ctx->pos_version = version;
    ctx->v_version++;
    if (A.write_force) ctx->fvalid[A.local_slot] = version;
    *consumed = (int)(q - k);
    return B2_OK;
}

static int run_one_step(b2_context* ctx) {
    const int T = 256;
    const int lo = ctx->a_lo, hi = ctx->a_hi, n = hi - lo, ndof = 3*n;   // owned range
    cudaStream_t s = ctx->stream;
    // the step counter of the RNG streams is advanced by the first kernel of the step when that is the
    // scalar prologue, otherwise by a kernel of its own
    phase_mark(ctx, B2_PHASE_INTEGRATE);
    const bool run_prologue = !ctx->ops.empty() && ctx->ops[0].kind == B2_OP_GLOBAL &&
                              !(ctx->ops[0].d == 1 && ctx->prologue_valid);
    const bool folded_begin = run_prologue;
    if (!folded_begin && ctx->uses_random) {
        k_step_begin<<<1, 1, 0, s>>>(ctx->rng_state);
        B2_LAUNCH_CHECK();
    }
    // the configuration at the start of a step satisfies the constraints: it is the reference for the
    // first position constraint of the step (OpenMM keeps `oldPos` the same way)
    for (const b2_op& op : ctx->ops)
        if (op.kind == B2_OP_CONSTRAIN_X) { B2_TRY(con_snapshot(ctx)); break; }
    for (size_t k = 0; k < ctx->ops.size(); k++) {
        const b2_op& op = ctx->ops[k];
        if (k > 0 && ctx->ops[k-1].kind != B2_OP_MVV_FACTOR) ctx->mvv_factor_hint = -1;   // a hint lives for one op
        switch (op.kind) {
        case B2_OP_EVAL: {
            static const bool dual_allowed = getenv("B2_NO_DUAL") == nullptr;
            const bool stale = ctx->fbuf[op.b] && ctx->fvalid[op.b] != ctx->pos_version;
            int partner = -1;
            if (dual_allowed && stale && !ctx->profiling && has_pair_force(ctx, (uint32_t)op.a)) {
                // a later evaluation at the same positions that also needs pair work?
                for (size_t q = k + 1; q < ctx->ops.size() && ctx->ops[q].kind == B2_OP_EVAL; q++) {
                    const b2_op& o = ctx->ops[q];
                    if (o.b != op.b && ctx->fbuf[o.b] && ctx->fvalid[o.b] != ctx->pos_version &&
                        has_pair_force(ctx, (uint32_t)o.a) && !((uint32_t)o.a & (uint32_t)op.a)) {
                        partner = (int)q;
                        break;
                    }
                }
            }
            if (partner >= 0)
                B2_TRY(forces_ensure_dual(ctx, (uint32_t)op.a, op.b, (uint32_t)ctx->ops[partner].a, ctx->ops[partner].b));
            else
                B2_TRY(forces_ensure(ctx, (uint32_t)op.a, op.b));
            break;
        }
        case B2_OP_PERDOF: {
            PerDofTable tab = make_table(ctx);
            if (op.a == 0) B2_TRY(dist_before_move(ctx));
            if (k < ctx->jit_fn.size() && ctx->jit_fn[k]) {
                // the step's own kernel, compiled from its bytecode at program load (jit.cu)
                int dof_lo = 3*lo, dof_hi = 3*hi, target = op.a, serial = op.e;
                const double* consts = ctx->consts; const double* globals = ctx->globals;
                const unsigned long long* rng_state = ctx->rng_state;
                void* args[] = {&dof_lo, &dof_hi, &tab, &target, &consts, &globals, &rng_state, &serial};
                B2_TRY(jit_launch(ctx, (int)k, (unsigned)std::max(1, (ndof + T - 1)/T), T, args));
            } else {
                k_perdof<<<std::max(1, (ndof + T - 1)/T), T, 0, s>>>(3*lo, 3*hi, tab, op.a, ctx->code + op.b, op.c, ctx->consts,
                                                          ctx->globals, ctx->rng_state, op.e, 0);
                B2_LAUNCH_CHECK();
            }
            if (op.a == 0) ctx->pos_version++;
            if (op.a == 1) ctx->v_version++;
            break;
        }
        case B2_OP_SUM: {
            const int blocks = 296;
            if (op.d == 1) {
                k_mvv_partial<<<blocks, T, 0, s>>>(lo, hi, ctx->v, ctx->massd, ctx->sum_partial);
            } else {
                PerDofTable tab = make_table(ctx);
                if (k < ctx->jit_fn.size() && ctx->jit_fn[k]) {
                    int dof_lo = 3*lo, dof_hi = 3*hi;
                    const double* consts = ctx->consts; const double* globals = ctx->globals;
                    double* partial = ctx->sum_partial;
                    void* args[] = {&dof_lo, &dof_hi, &tab, &consts, &globals, &partial};
                    B2_TRY(jit_launch(ctx, (int)k, blocks, T, args));
                    ctx->counters[0]--;          // counted once, by the check below
                } else {
                    k_sum_partial<<<blocks, T, 0, s>>>(3*lo, 3*hi, tab, ctx->code + op.b, op.c, ctx->consts, ctx->globals,
                                                        ctx->sum_partial);
                }
            }
            B2_LAUNCH_CHECK();
            k_sum_final<<<1, 256, 0, s>>>(blocks, ctx->sum_partial, ctx->globals, op.a);
            B2_LAUNCH_CHECK();
            B2_TRY(dist_reduce_value(ctx, ctx->globals + op.a));
            break;
        }
        case B2_OP_GLOBAL:
            if (k == 0 && !run_prologue) break;          // invariant coefficients already in place
            k_global<<<1, 64, 0, s>>>(ctx->code + op.b, op.c, ctx->consts, ctx->nconsts, ctx->globals,
                                       ctx->nglobals, ctx->rng_state, ctx->d_energy, (k == 0 && folded_begin) ? 1 : 0);
            B2_LAUNCH_CHECK();
            if (k == 0 && op.d == 1) ctx->prologue_valid = true;
            break;
        case B2_OP_KICK: {
            int consumed = 0;
            B2_TRY(try_fused_run(ctx, k, &consumed));
            if (consumed > 0) { k += consumed - 1; break; }
            B2_TRY(launch_vel(ctx, op));
            break;
        }
        case B2_OP_DRIFT:
            B2_TRY(dist_before_move(ctx));
            k_drift<<<std::max(1, (ndof + T - 1)/T), T, 0, s>>>(3*lo, 3*hi, ctx->x, ctx->v, ctx->massd, ctx->globals, op.a);
            B2_LAUNCH_CHECK();
            ctx->pos_version++;
            break;
        case B2_OP_SCALE:
            k_scale<<<std::max(1, (ndof + T - 1)/T), T, 0, s>>>(3*lo, 3*hi, ctx->v, ctx->globals, op.a);
            B2_LAUNCH_CHECK();
            ctx->v_version++;
            break;
        case B2_OP_UPDATE_STATE:
            break;
        case B2_OP_MVV_FACTOR:
            ctx->mvv_factor_hint = (op.a == ctx->mvv_index) ? op.b : -1;
            break;
        case B2_OP_CONSTRAIN_X:
            B2_TRY(dist_before_move(ctx));
            B2_TRY(con_positions(ctx));
            break;
        case B2_OP_CONSTRAIN_V:
            B2_TRY(con_velocities(ctx));
            ctx->v_version++;
            break;
        case B2_OP_INVALIDATE:
            for (int g = 0; g < B2_FSLOTS; g++) ctx->fvalid[g] = -1;
            ctx->deriv_version = -1;
            break;
        case B2_OP_ENERGY: {
            if (ctx->deriv_version == ctx->pos_version) break;      // nothing moved since the last evaluation
            ctx->deriv_version = ctx->pos_version;
            B2_CUDA(cudaMemsetAsync(ctx->d_energy + 64, 0, 2*sizeof(double), s));
            bool prepared = false;
            for (const PairForce& pf : ctx->pair_forces) {
                if (pf.family != B2_PAIR_SOFTCORE) continue;
                if (!prepared) {
                    if (!ctx->p2p) B2_TRY(dist_sync_positions(ctx));
                    B2_TRY(nl_prepare(ctx, false));
                    prepared = true;
                }
                B2_TRY(pair_eval_energy(ctx, pf, pf.group));
                B2_TRY(dist_allreduce(ctx, ctx->d_energy + 72, 4));
                k_fold_derivatives<<<1, 1, 0, s>>>(ctx->d_energy);
                B2_LAUNCH_CHECK();
            }
            break;
        }
        default:
            return b2_fail(ctx, B2_ERR_UNSUPPORTED, "unknown program op %d", op.kind);
        }
    }
    B2_TRY(dist_before_move(ctx));      // a step leaves no acknowledgement outstanding (graph replays are self-contained)
    return B2_OK;
}

// 0.5 sum m v.v over all atoms (report cadence: State.getKineticEnergy).  Fixed reduction tree over the owned range,
// rank-ordered sum across ranks: the same bits on every rank and in every run.
extern "C" int b2_kinetic_energy(b2_context* ctx, double* out_host) {
    if (!ctx || !out_host) return B2_ERR_ARG;
    if (!ctx->have_order) return b2_fail(ctx, B2_ERR_STATE, "positions have not been set");
    const int blocks = 296, T = 256;
    B2_TRY(ensure_partials(ctx, blocks + 1));
    k_mvv_partial<<<blocks, T, 0, ctx->stream>>>(ctx->a_lo, ctx->a_hi, ctx->v, ctx->massd, ctx->sum_partial);
    B2_LAUNCH_CHECK();
    double* slot = ctx->d_energy + 81;
    k_sum_final<<<1, 256, 0, ctx->stream>>>(blocks, ctx->sum_partial, slot, 0);
    B2_LAUNCH_CHECK();
    B2_TRY(dist_reduce_value(ctx, slot));
    double mvv = 0;
    B2_CUDA(cudaMemcpyAsync(&mvv, slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    *out_host = 0.5*mvv;
    return B2_OK;
}

int program_release(b2_context* ctx) {
    if (ctx->graph_exec) {
        cudaStreamSynchronize(ctx->stream);      // b2_run is asynchronous: the graph may still be in flight
        cudaGraphExecDestroy(ctx->graph_exec);
        ctx->graph_exec = nullptr;
    }
    ctx->graph_ready = false;
    ctx->eager_steps = 0;      // one eager step restores the steady-state force validity before the capture
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
    B2_TRY(inner_prepare(ctx));
    B2_TRY(con_prepare(ctx));
    B2_TRY(jit_prepare(ctx));            // generic per-DOF / sum steps become kernels of their own (once per program)
    B2_TRY(ensure_partials(ctx, (ctx->a_hi - ctx->a_lo + 255)/256 + 1));
    static const bool graph_allowed = getenv("B2_NO_GRAPH") == nullptr;
    const bool use_graph = graph_allowed && !ctx->profiling;
    static const int order_period = getenv("B2_ORDER_PERIOD") ? atoi(getenv("B2_ORDER_PERIOD")) : 250;
    bool has_update_state = false;       // the step program contains the UpdateContextState hook
    for (const b2_op& op : ctx->ops)
        if (op.kind == B2_OP_UPDATE_STATE) has_update_state = true;
    for (int done = 0; done < nsteps; done++) {
        if (order_period > 0 && ++ctx->steps_since_order_check >= order_period) {
            ctx->steps_since_order_check = 0;
            static const bool timing = getenv("B2_DEBUG_TIMING") != nullptr;
            const auto t0 = std::chrono::steady_clock::now();
            B2_TRY(order_refresh(ctx));          // may re-sort: chunks, clusters and the graph are rebuilt
            const auto t1 = std::chrono::steady_clock::now();
            B2_TRY(inner_prepare(ctx));
            B2_TRY(con_prepare(ctx));
            if (timing)
                fprintf(stderr, "[b2 order check] refresh %.2f ms, chunk/cluster tables %.2f ms\n",
                        std::chrono::duration<double, std::milli>(t1 - t0).count(),
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
        }
        B2_TRY(dist_before_move(ctx));
        if (ctx->baro_on && has_update_state && ++ctx->baro_steps >= ctx->baro_frequency) {
            ctx->baro_steps = 0;
            B2_TRY(barostat_attempt(ctx));           // may release the graph (accepted move: new box)
        }
        const unsigned long long entry = valid_mask(ctx);
        const bool synced = ctx->x_synced == ctx->pos_version;
        const bool mvv_ok = ctx->mvv_index >= 0 && ctx->mvv_version == ctx->v_version;
        if (use_graph && ctx->graph_ready && (entry & ctx->graph_entry_mask) == ctx->graph_entry_mask &&
            (synced || !ctx->graph_entry_synced) && (mvv_ok || !ctx->graph_entry_mvv)) {
            B2_CUDA(cudaGraphLaunch(ctx->graph_exec, ctx->stream));
            ctx->counters[5]++;
            ctx->v_version += ctx->graph_dv;
            if (ctx->graph_exit_mvv) ctx->mvv_version = ctx->v_version;
            ctx->pos_version += ctx->graph_dpos;
            if (ctx->graph_exit_synced) ctx->x_synced = ctx->pos_version;
            for (int g = 0; g < B2_FSLOTS; g++)
                if (ctx->graph_exit_mask & (1ull << g)) ctx->fvalid[g] = ctx->pos_version;
            continue;
        }
        if (use_graph && !ctx->graph_ready && ctx->eager_steps >= 1) {
            const long long v0 = ctx->pos_version;
            const long long vv0 = ctx->v_version;
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
            ctx->graph_entry_mvv = mvv_ok;
            ctx->graph_exit_mvv = ctx->mvv_index >= 0 && ctx->mvv_version == ctx->v_version;
            ctx->graph_dv = ctx->v_version - vv0;
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

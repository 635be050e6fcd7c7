// Explicit-list interactions (kernel K4 of SURVEY 2.1): harmonic bonds and angles, periodic
// torsions, LJ+Coulomb exception pairs and arbitrary CustomBondForce / CustomAngleForce energies.
//
// Replaces OpenMM's HarmonicBondForce / HarmonicAngleForce / PeriodicTorsionForce and the
// CustomBondForce objects atomsmm builds (forces.py:400-407 NonbondedExceptionsForce,
// forces.py:673-680 NearExceptionForce, systems.py:168,228 special bond/angle,
// systems.py:913-932 virial expressions).  One thread per term, geometry in float64 from the
// unwrapped master positions, forces accumulated in 64-bit fixed point (order-independent) and then
// added to the group's fp32 force buffer.
#include <math.h>

#include "bonded.cuh"

#define FULL 0xffffffffu

template <bool FORCE, bool ENERGY>
__global__ void k_bonded(BondArgs a, int arity, const double* x, unsigned long long* facc, double* acc) {
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    double e = 0, w = 0;
    if (t < a.nterms) {
        const GlobalGeo geo{x, facc};
        if (arity == 2) term_bond2<FORCE, ENERGY>(a, t, geo, e, w);
        else if (arity == 3) term_angle<FORCE, ENERGY>(a, t, geo, e);
        else term_torsion<FORCE, ENERGY>(a, t, geo, e);
    }
    if (ENERGY) block_accumulate(e, w, acc);
}

__global__ void k_bonded64(BondArgs a, int arity, const double* x, double* out) {
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= a.nterms) return;
    double e = 0, w = 0;
    const GlobalGeo64 geo{x, out};
    if (arity == 2) term_bond2<true, false>(a, t, geo, e, w);
    else if (arity == 3) term_angle<true, false>(a, t, geo, e);
    else term_torsion<true, false>(a, t, geo, e);
}

// out += fixed-point sums; the accumulators are left zeroed for the next use
__global__ void k_fixed_to_force(int lo, int hi, unsigned long long* __restrict__ facc, float4* __restrict__ out) {
    const int i = lo + blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const long long ax = (long long)facc[3*(size_t)i], ay = (long long)facc[3*(size_t)i+1], az = (long long)facc[3*(size_t)i+2];
    if ((ax | ay | az) == 0) return;
    float4 f = out[i];
    f.x += (float)((double)ax*(1.0/B2_FIXED_SCALE));
    f.y += (float)((double)ay*(1.0/B2_FIXED_SCALE));
    f.z += (float)((double)az*(1.0/B2_FIXED_SCALE));
    out[i] = f;
    facc[3*(size_t)i] = 0ull; facc[3*(size_t)i+1] = 0ull; facc[3*(size_t)i+2] = 0ull;
}

static int ensure_bond_acc(b2_context* ctx) {
    if (ctx->bond_acc) return B2_OK;
    B2_CUDA(cudaMalloc(&ctx->bond_acc, sizeof(unsigned long long)*3*(size_t)ctx->n));
    B2_CUDA(cudaMemsetAsync(ctx->bond_acc, 0, sizeof(unsigned long long)*3*(size_t)ctx->n, ctx->stream));
    return B2_OK;
}

static int flush_bond_acc(b2_context* ctx, float4* out) {
    // terms belong to the owner of their first atom and are intramolecular: only owned atoms receive force
    const int T = 256, lo = ctx->a_lo, hi = ctx->a_hi;
    k_fixed_to_force<<<std::max(1, (hi - lo + T - 1)/T), T, 0, ctx->stream>>>(lo, hi, ctx->bond_acc, out);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

__global__ void k_bonded_batch(BondBatch b, const double* x, unsigned long long* facc) {
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= b.first[b.count]) return;
    int k = 0;
    while (t >= b.first[k+1]) k++;
    double e = 0, w = 0;
    const int local = t - b.first[k];
    const GlobalGeo geo{x, facc};
    if (b.arity[k] == 2) term_bond2<true, false>(b.a[k], local, geo, e, w);
    else if (b.arity[k] == 3) term_angle<true, false>(b.a[k], local, geo, e);
    else term_torsion<true, false>(b.a[k], local, geo, e);
}

template <bool FORCE, bool ENERGY>
static int launch(b2_context* ctx, const BondedForce& bf, const BondArgs& a, float4* out, double* acc) {
    const int T = 128, blocks = (bf.nterms + T - 1)/T;
    if (FORCE) B2_TRY(ensure_bond_acc(ctx));
    k_bonded<FORCE, ENERGY><<<blocks, T, 0, ctx->stream>>>(a, bf.arity, ctx->x, ctx->bond_acc, acc);
    B2_LAUNCH_CHECK();
    if (FORCE) B2_TRY(flush_bond_acc(ctx, out));
    return B2_OK;
}

BondArgs bonded_make_args(b2_context* ctx, const BondedForce& bf) {
    BondArgs a;
    a.nterms = bf.nterms; a.stride = bf.stride; a.periodic = bf.periodic && ctx->periodic; a.family = bf.family;
    a.atoms = bf.atoms; a.params = bf.params; a.inv = ctx->inv;
    a.a_lo = ctx->a_lo; a.a_hi = ctx->a_hi;
    for (int k = 0; k < 3; k++) a.box[k] = ctx->box[k];
    for (int k = 0; k < 8; k++) a.g[k] = bf.gparams[k];
    a.code_e = bf.code_e; a.ncode_e = bf.ncode_e; a.code_de = bf.code_de; a.ncode_de = bf.ncode_de;
    a.consts = bf.consts;
    return a;
}

// forces of every explicit-list force whose group is in `mask`, accumulated into `out`
int bonded_eval_forces(b2_context* ctx, uint32_t mask, float4* out) {
    BondBatch b;
    b.count = 0;
    b.first[0] = 0;
    bool any = false;
    auto flush = [&]() -> int {
        if (b.count == 0) return B2_OK;
        B2_TRY(ensure_bond_acc(ctx));
        const int T = 128, total = b.first[b.count];
        k_bonded_batch<<<(total + T - 1)/T, T, 0, ctx->stream>>>(b, ctx->x, ctx->bond_acc);
        B2_LAUNCH_CHECK();
        b.count = 0;
        any = true;
        return B2_OK;
    };
    for (const BondedForce& bf : ctx->bonded_forces) {
        if (!(mask & (1u << bf.group)) || bf.nterms == 0) continue;
        if ((bf.family == B2_BOND_CUSTOM || bf.family == B2_ANGLE_CUSTOM) && bf.ncode_de == 0) continue;
        if (b.count == B2_MAX_BATCH) B2_TRY(flush());
        b.a[b.count] = bonded_make_args(ctx, bf);
        b.arity[b.count] = bf.arity;
        b.first[b.count + 1] = b.first[b.count] + bf.nterms;
        b.count++;
    }
    B2_TRY(flush());
    if (any) B2_TRY(flush_bond_acc(ctx, out));
    return B2_OK;
}

// accumulates forces into `out` (atomics) and, if want_energy, e / virial into d_energy+72..73
int bonded_eval(b2_context* ctx, const BondedForce& bf, float4* out, bool want_force, bool want_energy) {
    if (bf.nterms == 0) return B2_OK;
    BondArgs a = bonded_make_args(ctx, bf);
    double* acc = ctx->d_energy + 72;
    if (want_energy) B2_CUDA(cudaMemsetAsync(acc, 0, 4*sizeof(double), ctx->stream));
    if (want_force && want_energy) return launch<true, true>(ctx, bf, a, out, acc);
    if (want_force) return launch<true, false>(ctx, bf, a, out, acc);
    return launch<false, true>(ctx, bf, a, out, acc);
}

// float64 forces of every explicit-list force whose group is in `mask`, accumulated into out[n][3]
int bonded_eval_forces64(b2_context* ctx, uint32_t mask, double* out) {
    for (const BondedForce& bf : ctx->bonded_forces) {
        if (!(mask & (1u << bf.group)) || bf.nterms == 0) continue;
        if ((bf.family == B2_BOND_CUSTOM || bf.family == B2_ANGLE_CUSTOM) && bf.ncode_de == 0) continue;
        const int T = 128;
        k_bonded64<<<(bf.nterms + T - 1)/T, T, 0, ctx->stream>>>(bonded_make_args(ctx, bf), bf.arity, ctx->x, out);
        B2_LAUNCH_CHECK();
    }
    return B2_OK;
}

// Explicit-list interactions (kernel K4 of SURVEY 2.1): harmonic bonds and angles, periodic
// torsions, LJ+Coulomb exception pairs and arbitrary CustomBondForce / CustomAngleForce energies.
//
// Replaces OpenMM's HarmonicBondForce / HarmonicAngleForce / PeriodicTorsionForce and the
// CustomBondForce objects atomsmm builds (forces.py:400-407 NonbondedExceptionsForce,
// forces.py:673-680 NearExceptionForce, systems.py:168,228 special bond/angle,
// systems.py:913-932 virial expressions).  One thread per term, geometry in float64 from the
// unwrapped master positions, fp32 atomics into the group's force buffer.
#include <math.h>

#include "bonded.cuh"

#define FULL 0xffffffffu

template <bool FORCE, bool ENERGY>
__global__ void k_bonded(BondArgs a, int arity, const double* x, float4* out, double* acc) {
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    double e = 0, w = 0;
    if (t < a.nterms) {
        const GlobalGeo geo{x, out};
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

__global__ void k_bonded_batch(BondBatch b, const double* x, float4* out) {
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= b.first[b.count]) return;
    int k = 0;
    while (t >= b.first[k+1]) k++;
    double e = 0, w = 0;
    const int local = t - b.first[k];
    const GlobalGeo geo{x, out};
    if (b.arity[k] == 2) term_bond2<true, false>(b.a[k], local, geo, e, w);
    else if (b.arity[k] == 3) term_angle<true, false>(b.a[k], local, geo, e);
    else term_torsion<true, false>(b.a[k], local, geo, e);
}

template <bool FORCE, bool ENERGY>
static int launch(b2_context* ctx, const BondedForce& bf, const BondArgs& a, float4* out, double* acc) {
    const int T = 128, blocks = (bf.nterms + T - 1)/T;
    k_bonded<FORCE, ENERGY><<<blocks, T, 0, ctx->stream>>>(a, bf.arity, ctx->x, out, acc);
    B2_LAUNCH_CHECK();
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
    auto flush = [&]() -> int {
        if (b.count == 0) return B2_OK;
        const int T = 128, total = b.first[b.count];
        k_bonded_batch<<<(total + T - 1)/T, T, 0, ctx->stream>>>(b, ctx->x, out);
        B2_LAUNCH_CHECK();
        b.count = 0;
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
    return flush();
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

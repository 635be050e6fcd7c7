// Pair-tile kernels (K3 of SURVEY 2.1): forces in fp32 on the hot path, energies / virials /
// dE/dlambda in fp64 at report cadence, and the exact interacting-pair set for parity tests.
//
// Replaces the pair loops OpenMM runs for CustomNonbondedForce / NonbondedForce direct space
// (reference call sites: forces.py:225,153; evaluation through Context.getState, utils.py:164).
//
// Mapping: one warp per i-group of 8 spatially adjacent atoms; lane = (i-atom 0..7) x (j-lane
// 0..3).  The group's j-list is streamed 32 entries at a time: every lane gathers one j atom
// (fixed-point position + float4 parameters, coalesced index read), forms its minimum-image position
// relative to the group's first atom (integer wrap-around), converts to fp32 and stages it in shared
// memory; the warp then sweeps the 32 staged atoms in
// 8 steps of 4, each lane accumulating the force on its own i-atom in registers.  Forces are
// reduced over the 4 j-lanes with two shuffles and written once per atom: no atomics, results
// are bit-reproducible.  The full (both-directions) list means every pair is evaluated twice;
// that trades arithmetic for zero scatter traffic.
#include <math.h>

#include <algorithm>

#include "ctx.h"
#include "potentials.cuh"

#define FULL 0xffffffffu
#define WPB 4   // warps per block
// tuning knobs of the packed tiles (measured on B200: profiles/round2_pair_variants.txt)
#ifndef B2_PAIR_MINB
#define B2_PAIR_MINB 5          // resident blocks per SM the register budget is sized for
#endif
#ifndef B2_PAIR_UNROLL
#define B2_PAIR_UNROLL 4        // steps of a chunk the scheduler may interleave
#endif

template <typename T>
static PotParams<T> make_params(const b2_context* ctx, const PairForce& pf, float* rc2_out) {
    PotParams<T> p;
    memset(&p, 0, sizeof(p));
    p.sign = T(1);
    p.degree = 1;
    double rc_eff = pf.cutoff;
    const double* a = pf.params;
    auto set_switch = [&](bool use, double rs, double rc) {
        p.rs = T(rs);
        p.iw = (use && rc > rs) ? T(1.0/(rc - rs)) : T(0);
    };
    switch (pf.family) {
    case B2_PAIR_NEAR: {
        // variant, rs0, rc0, Kc, sign, use_coulomb
        const double rs = a[1], rc = a[2];
        set_switch(true, rs, rc);
        p.kc = T(a[5] != 0.0 ? a[3] : 0.0);
        p.sign = T(a[4]);
        p.inv_rc0 = T(1.0/rc);
        const double b = rs/(rc - rs);
        p.b = T(b);
        p.f12c = T(pow(1 + b, 3)*(pow(b, 6) + 3*pow(b, 5) + (30.0/7)*pow(b, 4) + (25.0/7)*pow(b, 3)
                                   + (25.0/14)*b*b + 0.5*b + 2.0/33)/pow(b, 9));
        p.f6c = T(pow(1 + b, 3)/pow(b, 3));
        p.f1c = T((30*(1 + b))*(b*b*(1 + b)*(1 + b)*log(1/b + 1) - b*b*b - 1.5*b*b - b/3 + 1.0/12));
        rc_eff = std::min(rc_eff, rc);
        break;
    }
    case B2_PAIR_DAMPED: {
        // alpha, rswitch, rcut, degree, Kc
        const double rs = a[1], rc = a[2];
        const int d = (int)a[3];
        set_switch(true, rs, rc);
        p.degree = d;
        p.rsd = T(pow(rs, d));
        p.iwd = T(1.0/(pow(rc, d) - pow(rs, d)));
        p.alpha = T(a[0]);
        p.tasp = T(2.0*a[0]/sqrt(M_PI));
        p.kc = T(a[4]);
        rc_eff = std::min(rc_eff, rc);
        break;
    }
    case B2_PAIR_LJC: {
        // Kc, coulomb_kind, krf, crf, alpha, use_switch, rswitch, rcut
        p.kc = T(a[1] != 0.0 ? a[0] : 0.0);
        p.krf = T(a[2]); p.crf = T(a[3]);
        p.alpha = T(a[4]); p.tasp = T(2.0*a[4]/sqrt(M_PI));
        set_switch(a[5] != 0.0, a[6], a[7]);
        break;
    }
    case B2_PAIR_LJ_VIRIAL: {
        set_switch(a[0] != 0.0, a[1], a[2]);
        break;
    }
    case B2_PAIR_SOFTCORE: {
        // Kc, lambda_vdw, lambda_coul, use_switch, rswitch, rcut
        p.kc = T(a[0]); p.lam_v = T(a[1]); p.lam_c = T(a[2]);
        set_switch(a[3] != 0.0, a[4], a[5]);
        p.gmode = T(pf.nparams > 6 ? a[6] : 0.0);
        if (ctx->globals && pf.bind[1] >= 0) p.lam_v_dev = ctx->globals + pf.bind[1];
        if (ctx->globals && pf.bind[2] >= 0) p.lam_c_dev = ctx->globals + pf.bind[2];
        break;
    }
    }
    *rc2_out = (float)(rc_eff*rc_eff);
    return p;
}

// ---------------------------------------------------------------------------------------------
// fp32 force kernel
// ---------------------------------------------------------------------------------------------
// Pairs whose fp32 r^2 lies within rounding distance of rc^2 (a few hundred per launch) are not
// decided here: they are recorded and settled by k_pair_band in float64 from the master
// coordinates, with the oracle's operation order, so that a pair is in or out identically in both
// atoms' tiles and identically to the float64 reference (Newton's third law holds exactly even for
// potentials that are discontinuous at the cutoff, e.g. reaction field).
struct BandBuffer {
    int* pairs;          // int2 (sorted i, sorted j)
    unsigned* count;
    unsigned capacity;
    int* slot;           // [n] per atom: index of the band entry that owns the atom's accumulator, or -1
    long long* acc;      // [capacity][3] fixed-point accumulators (2^-32 kJ/mol/nm)
    unsigned* ticket;    // last-block-done counter of k_pair_band
    int* flags;          // nl_flags: [10] = band overflow
};

template <class POT, bool MINIMG>
__device__ __forceinline__ void sweep_chunk(const POT& pot, const float4* __restrict__ sx,
                                            const float4* __restrict__ sp, int jj, int il, float4 xi,
                                            float qi, float hsi, float sei, float rc2, float3 box, float3 inv,
                                            int i, float rc2_lo, const BandBuffer& bb,
                                            double& ax, double& ay, double& az) {
    float fx = 0.f, fy = 0.f, fz = 0.f;   // fp32 partial sums over one chunk (<= 8 pairs per lane)
#pragma unroll
    for (int s = 0; s < 8; s++) {
        const int jl = 4*s + jj;
        const float4 xj = sx[jl];
        const float4 pj = sp[jl];
        float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
        if (MINIMG) {
            dx -= box.x*rintf(dx*inv.x);
            dy -= box.y*rintf(dy*inv.y);
            dz -= box.z*rintf(dz*inv.z);
        }
        const float r2 = dx*dx + dy*dy + dz*dz;
        const unsigned en = (unsigned)__float_as_int(pj.w);      // full list entry: mask<<24 | j
        if (r2 < rc2 && !((en >> (24 + il)) & 1u)) {      // rc2 here is the OUTER edge of the band
            if (r2 >= rc2_lo) {
                if (i >= 0) {                     // padding lanes of the last group own no atom
                    const unsigned slot = atomicAdd(bb.count, 1u);
                    if (slot < bb.capacity) { bb.pairs[2*slot] = i; bb.pairs[2*slot+1] = (int)(en & 0xffffffu); }
                }
                continue;
            }
            float rF, e, rinv2;
            pot.template operator()<false>(r2, qi*pj.x, hsi + pj.y, sei*pj.z, rF, e, rinv2);
            const float fr = rF*rinv2;
            fx += fr*dx; fy += fr*dy; fz += fr*dz;
        }
    }
    ax += (double)fx; ay += (double)fy; az += (double)fz;   // fp64 across chunks: no long fp32 sums
}

// Positions reach the tiles as 32-bit FIXED-POINT fractions of the box (xq, maintained by the skin-test
// kernel): one 16-byte load per staged atom instead of three strided doubles, and the difference of two
// fixed-point coordinates wraps around exactly like the periodic box does -- the minimum image relative to the
// group's reference atom costs nothing.  Resolution L/2^32 (8e-9 nm at L = 35 nm); the difference is exact,
// its conversion to fp32 has the error of a ~1 nm vector (6e-8 nm), independent of the box size.
__device__ __forceinline__ float3 rel_fixed(int4 q, int4 ref, float3 scale) {
    return make_float3((float)(q.x - ref.x)*scale.x, (float)(q.y - ref.y)*scale.y, (float)(q.z - ref.z)*scale.z);
}

// The list of a group is streamed 32 entries at a time through a THREE-stage pipeline: while chunk c is swept
// from shared memory, the gathers (fixed-point position, parameters) of chunk c+1 are in flight and so is the
// index load of chunk c+2 -- the dependent chain entries -> gathers never sits on the critical path, and the
// gathered values are first touched after the sweep.
template <class POT>
__global__ void __launch_bounds__(32*WPB, 8) k_pair_force(int n, int g_lo, int ngroups, const int4* __restrict__ xq,
                                                      const float4* __restrict__ par,
                                                      const int* __restrict__ entries,
                                                      const int* __restrict__ counts,
                                                      const unsigned char* __restrict__ gflags, int cap,
                                                      float4* __restrict__ out, int accumulate, POT pot,
                                                      float rc2, BandBuffer bb, float3 box) {
    __shared__ float4 sx[WPB][2][32];
    __shared__ float4 sp[WPB][2][32];
    const int warp = g_lo + ((blockIdx.x*blockDim.x + threadIdx.x) >> 5);
    if (warp >= ngroups) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int il = lane >> 2, jj = lane & 3;
    const int i = warp*B2_GROUP + il;
    const int ic = min(i, n - 1);
    const float3 inv = make_float3(1.f/box.x, 1.f/box.y, 1.f/box.z);
    const float3 scale = make_float3(box.x*2.3283064365386963e-10f, box.y*2.3283064365386963e-10f,
                                     box.z*2.3283064365386963e-10f);
    // reference point of the group: its first atom
    const int i0 = warp*B2_GROUP;
    const int4 ref = xq[i0];
    const float3 xr = rel_fixed(xq[ic], ref, scale);
    const float4 xi = make_float4(xr.x, xr.y, xr.z, 0.f);
    const float4 pi = par[ic];
    const bool minimg = gflags[warp] & 1;
    const float rc2_lo = rc2*(1.f - 2e-6f), rc2_hi = rc2*(1.f + 2e-6f);
    const int cnt = counts[warp];
    const int* __restrict__ base = entries + (size_t)warp*cap;
    const int pad = (int)(0xff000000u | (unsigned)i0);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    int buf = 0;
    int e = lane < cnt ? base[lane] : pad;                       // chunk 0
    int e_next = 32 + lane < cnt ? base[32 + lane] : pad;        // chunk 1
    int4 qj = xq[e & 0xffffff];
    float4 pj = par[e & 0xffffff];
    for (int c0 = 0; c0 < cnt; c0 += 32) {
        const float3 xj = rel_fixed(qj, ref, scale);
        pj.w = __int_as_float(e);
        sx[wib][buf][lane] = make_float4(xj.x, xj.y, xj.z, 0.f);
        sp[wib][buf][lane] = pj;
        if (c0 + 32 < cnt) {
            e = e_next;
            qj = xq[e & 0xffffff];
            pj = par[e & 0xffffff];
            const int nxt = c0 + 64 + lane;
            e_next = nxt < cnt ? base[nxt] : pad;
        }
        __syncwarp();
        if (minimg)
            sweep_chunk<POT, true>(pot, sx[wib][buf], sp[wib][buf], jj, il, xi, pi.x, pi.y, pi.z, rc2_hi, box, inv, i < n ? i : -1, rc2_lo, bb, fx, fy, fz);
        else
            sweep_chunk<POT, false>(pot, sx[wib][buf], sp[wib][buf], jj, il, xi, pi.x, pi.y, pi.z, rc2_hi, box, inv, i < n ? i : -1, rc2_lo, bb, fx, fy, fz);
        buf ^= 1;
    }
    fx += __shfl_xor_sync(FULL, fx, 1); fy += __shfl_xor_sync(FULL, fy, 1); fz += __shfl_xor_sync(FULL, fz, 1);
    fx += __shfl_xor_sync(FULL, fx, 2); fy += __shfl_xor_sync(FULL, fy, 2); fz += __shfl_xor_sync(FULL, fz, 2);
    if (jj == 0 && i < n) {
        float4 f = make_float4((float)fx, (float)fy, (float)fz, 0.f);
        if (accumulate) {
            const float4 o = out[i];
            f.x += o.x; f.y += o.y; f.z += o.z;
        }
        out[i] = f;
    }
}

// ---------------------------------------------------------------------------------------------
// Packed variant (B2_PAIR_PACKED=1; NOT the default, see launch_force): every lane evaluates TWO list slots per step
// with the f32x2 instructions of sm_100a (FADD2 / FMUL2 / FFMA2, potentials.cuh) -- one issue carries the arithmetic
// of two pairs.  Shared-memory staging is laid out for that: per 32-entry chunk 16 slot
// PAIRS, each {(x0,x1,y0,y1), (z0,z1,q0,q1), (hs0,hs1,se0,se1), (entry0,entry1)}, so an LDS.128 lands directly in
// the aligned register pairs the packed instructions take.  Pipeline, band logic and reduction as above.
// Differences are accumulated as x_j - x_i (one packed add against the negated i position) and the sign is
// applied once at the end.
__device__ __forceinline__ void band_record(const BandBuffer& bb, int i, unsigned entry) {
    if (i < 0) return;                                  // padding lanes of the last group own no atom
    const unsigned slot = atomicAdd(bb.count, 1u);
    if (slot < bb.capacity) { bb.pairs[2*slot] = i; bb.pairs[2*slot+1] = (int)(entry & 0xffffffu); }
}

// MASKED = false: no entry of the chunk carries exclusion bits (decided once per chunk by a warp vote at staging
// time; true for ~8 chunks in 10) -- the entry words are then read only by the rare band path.
// The four steps of a chunk are straight-line code: band pairs are only FLAGGED inside the loop and recorded after
// it, so that no branch separates the steps and their (long, dependent) arithmetic chains can be interleaved by
// the scheduler -- with ~6 warps per scheduler the tiles are otherwise bound by dependent-issue latency.
template <class POT2, bool MINIMG, bool MASKED>
__device__ __forceinline__ void sweep_chunk2(const POT2& pot, const float4* __restrict__ sA, const float4* __restrict__ sB,
                                             const float4* __restrict__ sC, const uint2* __restrict__ sD, int jj,
                                             unsigned ibit, F2 nx, F2 ny, F2 nz, F2 kqi, F2 hsi, F2 sei, float rc2_hi,
                                             float rc2_lo, float3 box, float3 inv, int i, const BandBuffer& bb,
                                             double& fx, double& fy, double& fz) {
    F2 ax = f2(0.f), ay = f2(0.f), az = f2(0.f);      // fp32 partial sums over one chunk (<= 8 pairs per lane)
    unsigned bandbits = 0u;
    constexpr int unroll = B2_PAIR_UNROLL;
#pragma unroll unroll
    for (int t = 0; t < 4; t++) {
        const int pp = 4*t + jj;
        const float4 A = sA[pp], B = sB[pp], C = sC[pp];
        uint2 D = make_uint2(0u, 0u);
        if (MASKED) D = sD[pp];
        F2 dx = f2(A.x, A.y) + nx, dy = f2(A.z, A.w) + ny, dz = f2(B.x, B.y) + nz;
        if (MINIMG) {
            dx.v.x -= box.x*rintf(dx.v.x*inv.x); dx.v.y -= box.x*rintf(dx.v.y*inv.x);
            dy.v.x -= box.y*rintf(dy.v.x*inv.y); dy.v.y -= box.y*rintf(dy.v.y*inv.y);
            dz.v.x -= box.z*rintf(dz.v.x*inv.z); dz.v.y -= box.z*rintf(dz.v.y*inv.z);
        }
        const F2 r2 = fma2(dz, dz, fma2(dy, dy, dx*dx));
        // hit: inside the INNER edge of the band and not excluded -> evaluated here; band: settled in float64
        const bool ex0 = D.x & ibit, ex1 = D.y & ibit;
        const bool hit0 = r2.v.x < rc2_lo && !ex0, hit1 = r2.v.y < rc2_lo && !ex1;
        const bool band0 = r2.v.x < rc2_hi && !ex0 && !hit0, band1 = r2.v.y < rc2_hi && !ex1 && !hit1;
        bandbits |= (band0 ? 1u << (2*t) : 0u) | (band1 ? 2u << (2*t) : 0u);
        F2 fr = pot(r2, kqi*f2(B.z, B.w), hsi + f2(C.x, C.y), sei*f2(C.z, C.w));
        fr = f2(hit0 ? fr.v.x : 0.f, hit1 ? fr.v.y : 0.f);    // select, not multiply: masked slots may hold inf / NaN
        ax = fma2(fr, dx, ax); ay = fma2(fr, dy, ay); az = fma2(fr, dz, az);
    }
    if (bandbits) {                                            // a few hundred pairs per launch
        for (int b = 0; b < 8; b++)
            if (bandbits >> b & 1u) band_record(bb, i, reinterpret_cast<const unsigned*>(&sD[4*(b >> 1) + jj])[b & 1]);
    }
    // d = x_j - x_i: the force on i is -sum.  fp64 across chunks: no long fp32 sums
    fx -= (double)(ax.v.x + ax.v.y); fy -= (double)(ay.v.x + ay.v.y); fz -= (double)(az.v.x + az.v.y);
}

template <class POT2>
__global__ void __launch_bounds__(32*WPB, B2_PAIR_MINB) k_pair_force2(int n, int g_lo, int ngroups, const int4* __restrict__ xq,
                                                       const float4* __restrict__ par,
                                                       const int* __restrict__ entries,
                                                       const int* __restrict__ counts,
                                                       const unsigned char* __restrict__ gflags, int cap,
                                                       float4* __restrict__ out, int accumulate, POT2 pot,
                                                       float rc2, BandBuffer bb, float3 box) {
    __shared__ float4 sA[WPB][2][16], sB[WPB][2][16], sC[WPB][2][16];
    __shared__ uint2 sD[WPB][2][16];
    const int warp = g_lo + ((blockIdx.x*blockDim.x + threadIdx.x) >> 5);
    if (warp >= ngroups) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int il = lane >> 2, jj = lane & 3;
    const int i = warp*B2_GROUP + il;
    const int ic = min(i, n - 1);
    const float3 inv = make_float3(1.f/box.x, 1.f/box.y, 1.f/box.z);
    const float3 scale = make_float3(box.x*2.3283064365386963e-10f, box.y*2.3283064365386963e-10f,
                                     box.z*2.3283064365386963e-10f);
    const int i0 = warp*B2_GROUP;
    const int4 ref = xq[i0];                      // reference point of the group: its first atom
    const float3 xr = rel_fixed(xq[ic], ref, scale);
    const float4 pi = par[ic];
    const F2 nx = f2(-xr.x), ny = f2(-xr.y), nz = f2(-xr.z);
    const F2 kqi = f2(pot.charge_scale()*pi.x), hsi = f2(pi.y), sei = f2(POT2::EPS_SCALE*pi.z);
    const unsigned ibit = 1u << (24 + il);
    const bool minimg = gflags[warp] & 1;
    const float rc2_lo = rc2*(1.f - 2e-6f), rc2_hi = rc2*(1.f + 2e-6f);
    const int cnt = counts[warp];
    const int* __restrict__ base = entries + (size_t)warp*cap;
    const int pad = (int)(0xff000000u | (unsigned)i0);
    const int half = lane & 1, pr = lane >> 1;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    int buf = 0;
    int e = lane < cnt ? base[lane] : pad;                       // chunk 0
    int e_next = 32 + lane < cnt ? base[32 + lane] : pad;        // chunk 1
    int4 qj = xq[e & 0xffffff];
    float4 pj = par[e & 0xffffff];
    for (int c0 = 0; c0 < cnt; c0 += 32) {
        const float3 xj = rel_fixed(qj, ref, scale);
        float* a = reinterpret_cast<float*>(&sA[wib][buf][pr]) + half;
        float* b = reinterpret_cast<float*>(&sB[wib][buf][pr]) + half;
        float* c = reinterpret_cast<float*>(&sC[wib][buf][pr]) + half;
        a[0] = xj.x; a[2] = xj.y;
        b[0] = xj.z; b[2] = pj.x;
        c[0] = pj.y; c[2] = pj.z;
        reinterpret_cast<int*>(&sD[wib][buf][pr])[half] = e;
        const bool masked = __any_sync(FULL, ((unsigned)e >> 24) != 0u);     // exclusions, self pairs, padding
        if (c0 + 32 < cnt) {
            e = e_next;
            qj = xq[e & 0xffffff];
            pj = par[e & 0xffffff];
            const int nxt = c0 + 64 + lane;
            e_next = nxt < cnt ? base[nxt] : pad;
        }
        __syncwarp();
        if (minimg)
            sweep_chunk2<POT2, true, true>(pot, sA[wib][buf], sB[wib][buf], sC[wib][buf], sD[wib][buf], jj, ibit, nx, ny, nz,
                                           kqi, hsi, sei, rc2_hi, rc2_lo, box, inv, i < n ? i : -1, bb, fx, fy, fz);
        else if (masked)
            sweep_chunk2<POT2, false, true>(pot, sA[wib][buf], sB[wib][buf], sC[wib][buf], sD[wib][buf], jj, ibit, nx, ny, nz,
                                            kqi, hsi, sei, rc2_hi, rc2_lo, box, inv, i < n ? i : -1, bb, fx, fy, fz);
        else
            sweep_chunk2<POT2, false, false>(pot, sA[wib][buf], sB[wib][buf], sC[wib][buf], sD[wib][buf], jj, ibit, nx, ny, nz,
                                             kqi, hsi, sei, rc2_hi, rc2_lo, box, inv, i < n ? i : -1, bb, fx, fy, fz);
        buf ^= 1;
    }
    fx += __shfl_xor_sync(FULL, fx, 1); fy += __shfl_xor_sync(FULL, fy, 1); fz += __shfl_xor_sync(FULL, fz, 1);
    fx += __shfl_xor_sync(FULL, fx, 2); fy += __shfl_xor_sync(FULL, fy, 2); fz += __shfl_xor_sync(FULL, fz, 2);
    if (jj == 0 && i < n) {
        float4 f = make_float4((float)fx, (float)fy, (float)fz, 0.f);
        if (accumulate) {
            const float4 o = out[i];
            f.x += o.x; f.y += o.y; f.z += o.z;
        }
        out[i] = f;
    }
}

// Settlement of the band pairs in float64.  The append order of the band buffer depends on warp
// scheduling, so the contributions are NOT added to the fp32 force buffer one by one (float addition
// does not commute with re-ordering): every touched atom gets ONE fixed-point accumulator (claimed
// by the first entry that reaches it; which entry wins only names the accumulator), all entries add
// into it with 64-bit integer atomics -- associative, hence independent of the order -- and a second
// pass (the last block of the same kernel, or a kernel of its own for large systems) adds every accumulator to
// the atom's force exactly once and cleans up.
#define B2_BAND_SCALE 4294967296.0
__device__ __forceinline__ void band_apply(const BandBuffer& bb, unsigned k, float4* __restrict__ out) {
    const int i = bb.pairs[2*k];
    if (__ldcg(&bb.slot[i]) != (int)k) return;                // not the entry that owns atom i's accumulator
    long long* a = bb.acc + 3*(size_t)k;
    const double fx = (double)__ldcg(&a[0])*(1.0/B2_BAND_SCALE), fy = (double)__ldcg(&a[1])*(1.0/B2_BAND_SCALE),
                 fz = (double)__ldcg(&a[2])*(1.0/B2_BAND_SCALE);
    float4 f = out[i];
    f.x += (float)fx; f.y += (float)fy; f.z += (float)fz;
    out[i] = f;
    a[0] = 0; a[1] = 0; a[2] = 0;
    bb.slot[i] = -1;
}

// fused_apply: the last block to finish also applies the accumulators (small systems: one launch instead of
// two); large systems launch k_pair_band_apply over many blocks instead
template <class POTD>
__global__ void k_pair_band(const double* __restrict__ x, const double* __restrict__ pard, BandBuffer bb, POTD pot,
                            double rc2d, double bx, double by, double bz, float4* __restrict__ out, int fused_apply) {
    const unsigned found = *bb.count;
    const unsigned total = min(found, bb.capacity);
    if (found > bb.capacity && blockIdx.x == 0 && threadIdx.x == 0) bb.flags[10] = 1;   // reported by b2_synchronize
    for (unsigned k = blockIdx.x*blockDim.x + threadIdx.x; k < total; k += gridDim.x*blockDim.x) {
        const int i = bb.pairs[2*k], j = bb.pairs[2*k+1];
        double dx = x[3*j] - x[3*i], dy = x[3*j+1] - x[3*i+1], dz = x[3*j+2] - x[3*i+2];
        dx -= bx*rint(dx/bx); dy -= by*rint(dy/by); dz -= bz*rint(dz/bz);
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (r2 < rc2d) {
            double rF, e, rinv2;
            pot.template operator()<false>(r2, pard[3*i]*pard[3*j], 0.5*(pard[3*i+1] + pard[3*j+1]),
                                           sqrt(pard[3*i+2]*pard[3*j+2]), rF, e, rinv2);
            const double fr = -rF*rinv2;        // d points from i to j
            const int prev = atomicCAS(&bb.slot[i], -1, (int)k);
            const int owner = prev < 0 ? (int)k : prev;
            unsigned long long* a = reinterpret_cast<unsigned long long*>(bb.acc + 3*(size_t)owner);
            atomicAdd(a, (unsigned long long)__double2ll_rn(fr*dx*B2_BAND_SCALE));
            atomicAdd(a + 1, (unsigned long long)__double2ll_rn(fr*dy*B2_BAND_SCALE));
            atomicAdd(a + 2, (unsigned long long)__double2ll_rn(fr*dz*B2_BAND_SCALE));
        }
    }
    if (!fused_apply) return;
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(bb.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (unsigned k = threadIdx.x; k < total; k += blockDim.x) band_apply(bb, k, out);
    if (threadIdx.x == 0) *bb.ticket = 0;
}

// second half of the settlement: every accumulator is added to its atom's force exactly once, then cleaned
__global__ void k_pair_band_apply(BandBuffer bb, float4* __restrict__ out) {
    const unsigned total = min(*bb.count, bb.capacity);
    for (unsigned k = blockIdx.x*blockDim.x + threadIdx.x; k < total; k += gridDim.x*blockDim.x) band_apply(bb, k, out);
}

// ---------------------------------------------------------------------------------------------
// fp64 energy / virial / dE/dlambda kernel (same list, per-pair minimum image, double state)
// ---------------------------------------------------------------------------------------------
template <class POT, bool SOFT>
__global__ void __launch_bounds__(32*WPB) k_pair_energy(int n, int g_lo, int ngroups, int a_lo, int a_hi,
                                                       const double* __restrict__ x,
                                                       const double* __restrict__ pard,
                                                       const int* __restrict__ entries,
                                                       const int* __restrict__ counts, int cap, POT pot,
                                                       double rc2, double bx, double by, double bz,
                                                       double* __restrict__ acc /* e, w, dlv, dlc */,
                                                       const int4* __restrict__ xq) {
    const int warp = g_lo + ((blockIdx.x*blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    double e_sum = 0, w_sum = 0, dv_sum = 0, dc_sum = 0;
    if (warp < ngroups) {
        const int il = lane >> 2, jj = lane & 3;
        const int i = warp*B2_GROUP + il;
        const int ic = min(i, n - 1);
        const bool mine = i >= a_lo && i < a_hi;
        const double xi = x[3*ic], yi = x[3*ic+1], zi = x[3*ic+2];
        const double qi = pard[3*ic], si = pard[3*ic+1], ei = pard[3*ic+2];
        const int cnt = counts[warp];
        const int* __restrict__ base = entries + (size_t)warp*cap;
        // fp32 PREFILTER on the fixed-point positions the force tiles use (one 16-byte gather, exact wrap-around
        // minimum image, relative error of r^2 ~ 2e-7): six in ten list slots are outside the cutoff and never reach
        // the float64 path (three strided double gathers, minimum image, exact criterion).  The margin of 1e-5 makes
        // the prefilter a strict superset of the float64 decision.
        const int4 qi4 = xq[ic];
        const float sxf = (float)(bx*2.3283064365386963e-10), syf = (float)(by*2.3283064365386963e-10),
                    szf = (float)(bz*2.3283064365386963e-10);
        const float rc2_pre = (float)rc2*(1.f + 1e-5f);
        const double ibx = 1.0/bx, iby = 1.0/by, ibz = 1.0/bz;
        for (int k = jj; k < cnt && mine; k += 4) {
            const int en = base[k];
            const int j = en & 0xffffff;
            if (((unsigned)en >> (24 + il)) & 1u) continue;
            const int4 qj4 = xq[j];
            const float fx = (float)(qi4.x - qj4.x)*sxf, fy = (float)(qi4.y - qj4.y)*syf, fz = (float)(qi4.z - qj4.z)*szf;
            if (fx*fx + fy*fy + fz*fz >= rc2_pre) continue;
            double dx = xi - x[3*j], dy = yi - x[3*j+1], dz = zi - x[3*j+2];
            dx -= bx*rint(dx*ibx); dy -= by*rint(dy*iby); dz -= bz*rint(dz*ibz);
            const double r2 = dx*dx + dy*dy + dz*dz;
            if (r2 < rc2) {
                double rF, e;
                const double qq = qi*pard[3*j], sig = 0.5*(si + pard[3*j+1]), eps = sqrt(ei*pard[3*j+2]);
                if constexpr (SOFT) {
                    double dv, dc;
                    pot.eval(r2, qq, sig, eps, rF, e, dv, dc);
                    dv_sum += dv; dc_sum += dc;
                } else {
                    double rinv2;
                    pot.template operator()<true>(r2, qq, sig, eps, rF, e, rinv2);
                }
                e_sum += e; w_sum += rF;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        e_sum += __shfl_xor_sync(FULL, e_sum, o);
        w_sum += __shfl_xor_sync(FULL, w_sum, o);
        if (SOFT) {
            dv_sum += __shfl_xor_sync(FULL, dv_sum, o);
            dc_sum += __shfl_xor_sync(FULL, dc_sum, o);
        }
    }
    if (lane == 0 && warp < ngroups) {
        // every pair appears in both atoms' lists
        atomicAdd(&acc[0], 0.5*e_sum);
        atomicAdd(&acc[1], 0.5*w_sum);
        if (SOFT) { atomicAdd(&acc[2], 0.5*dv_sum); atomicAdd(&acc[3], 0.5*dc_sum); }
    }
}

// ---------------------------------------------------------------------------------------------
// fp64 force kernel (report cadence: State forces with Precision=double, e.g. the molecular virial
// of computers.py:211-240, which contracts forces with positions and needs more than fp32)
// ---------------------------------------------------------------------------------------------
template <class POT>
__global__ void __launch_bounds__(32*WPB) k_pair_force64(int n, int g_lo, int ngroups, const double* __restrict__ x,
                                                        const double* __restrict__ pard,
                                                        const int* __restrict__ entries,
                                                        const int* __restrict__ counts, int cap, POT pot,
                                                        double rc2, double bx, double by, double bz,
                                                        double* __restrict__ out /* [n][3], accumulated */) {
    const int warp = g_lo + ((blockIdx.x*blockDim.x + threadIdx.x) >> 5);
    if (warp >= ngroups) return;
    const int lane = threadIdx.x & 31;
    const int il = lane >> 2, jj = lane & 3;
    const int i = warp*B2_GROUP + il;
    const int ic = min(i, n - 1);
    const double xi = x[3*ic], yi = x[3*ic+1], zi = x[3*ic+2];
    const double qi = pard[3*ic], si = pard[3*ic+1], ei = pard[3*ic+2];
    const int cnt = counts[warp];
    const int* __restrict__ base = entries + (size_t)warp*cap;
    double fx = 0, fy = 0, fz = 0;
    for (int k = jj; k < cnt; k += 4) {
        const int en = base[k];
        const int j = en & 0xffffff;
        if (((unsigned)en >> (24 + il)) & 1u) continue;
        double dx = xi - x[3*j], dy = yi - x[3*j+1], dz = zi - x[3*j+2];
        dx -= bx*rint(dx/bx); dy -= by*rint(dy/by); dz -= bz*rint(dz/bz);
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (r2 < rc2) {
            double rF, e, rinv2;
            pot.template operator()<false>(r2, qi*pard[3*j], 0.5*(si + pard[3*j+1]), sqrt(ei*pard[3*j+2]), rF, e, rinv2);
            const double fr = rF*rinv2;
            fx += fr*dx; fy += fr*dy; fz += fr*dz;
        }
    }
    fx += __shfl_xor_sync(FULL, fx, 1); fy += __shfl_xor_sync(FULL, fy, 1); fz += __shfl_xor_sync(FULL, fz, 1);
    fx += __shfl_xor_sync(FULL, fx, 2); fy += __shfl_xor_sync(FULL, fy, 2); fz += __shfl_xor_sync(FULL, fz, 2);
    if (jj == 0 && i < n) { out[3*i] += fx; out[3*i+1] += fy; out[3*i+2] += fz; }
}

// exact interacting pair set: i<j (caller numbering), r^2 < rc^2 in float64, not excluded
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30))*0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27))*0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(32*WPB) k_pair_set(int n, int ngroups, const double* __restrict__ x,
                                                    const int* __restrict__ orig,
                                                    const int* __restrict__ entries,
                                                    const int* __restrict__ counts, int cap, double rc2,
                                                    double bx, double by, double bz,
                                                    unsigned long long* __restrict__ acc, int* pairs,
                                                    long long capacity) {
    const int warp = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= ngroups) return;
    const int il = lane >> 2, jj = lane & 3;
    const int i = warp*B2_GROUP + il;
    if (i >= n) return;
    const double xi = x[3*i], yi = x[3*i+1], zi = x[3*i+2];
    const int oi = orig[i];
    const int cnt = counts[warp];
    const int* __restrict__ base = entries + (size_t)warp*cap;
    unsigned long long c = 0, h = 0;
    for (int k = jj; k < cnt; k += 4) {
        const int en = base[k];
        const int j = en & 0xffffff;
        if (((unsigned)en >> (24 + il)) & 1u) continue;
        const int oj = orig[j];
        if (oj <= oi) continue;
        double dx = x[3*j] - xi, dy = x[3*j+1] - yi, dz = x[3*j+2] - zi;
        dx -= bx*rint(dx/bx); dy -= by*rint(dy/by); dz -= bz*rint(dz/bz);
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (r2 < rc2) {
            c++;
            h += mix64(((unsigned long long)oi << 32) | (unsigned)oj);
            if (pairs) {
                unsigned long long slot = atomicAdd(&acc[2], 1ull);
                if ((long long)slot < capacity) { pairs[2*slot] = oi; pairs[2*slot+1] = oj; }
            }
        }
    }
    if (c) { atomicAdd(&acc[0], c); atomicAdd(&acc[1], h); }
}

// ---------------------------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------------------------
static double effective_cutoff(const PairForce& pf) {
    double rc = pf.cutoff;
    if (pf.family == B2_PAIR_NEAR || pf.family == B2_PAIR_DAMPED) rc = std::min(rc, pf.params[2]);
    return rc;
}

template <class POT>
struct PackedOf;
template <int A, int B, int C, int D, int E>
struct PackedOf<LJCPot<A, B, C, D, E, float>> { typedef LJCForce2<A, B, C, D, E> type; };
template <>
struct PackedOf<SoftcorePot<float>> { typedef SoftcoreForce2 type; };

template <class POT, class POTD>
static int launch_force(b2_context* ctx, const PairForce& pf, POT pot, POTD potd, float rc2, float4* out,
                        bool accumulate, int lane) {
    const double rcd = effective_cutoff(pf);
    // lane 1 = the side stream, with its own half of the band buffer
    cudaStream_t stream = lane ? ctx->side_stream : ctx->stream;
    BandBuffer bb{ctx->band_pairs + (lane ? 2*(size_t)ctx->band_capacity : 0), ctx->band_count + lane, ctx->band_capacity,
                  ctx->band_slot + (lane ? (size_t)ctx->n : 0), ctx->band_acc + (lane ? 3*(size_t)ctx->band_capacity : 0),
                  ctx->band_ticket + lane, ctx->nl_flags};
    B2_CUDA(cudaMemsetAsync(bb.count, 0, sizeof(unsigned), stream));
    const NList& L = ctx->lists[pf.list];
    const int blocks = (ctx->g_hi - ctx->g_lo + WPB - 1)/WPB;
    if (blocks == 0) return B2_OK;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->profiling) {
        cudaEventCreate(&ev0); cudaEventCreate(&ev1);
        cudaEventRecord(ev0, stream);
    }
    const float3 boxf = make_float3((float)ctx->box[0], (float)ctx->box[1], (float)ctx->box[2]);
    // B2_PAIR_PACKED=1 selects the f32x2 tiles (two list slots per lane and step).  Measured on B200
    // (profiles/round2_pair_variants.txt): FFMA2 / FMUL2 / FADD2 issue at HALF the rate of their scalar forms, so the
    // fp32 pipe sees the same work and the packed tiles are 4-13 % SLOWER than the one-slot-per-lane tiles in every
    // register / unrolling variant tried; they stay in the build as the measured alternative, not as the default.
    static const bool scalar_pair_tiles = getenv("B2_PAIR_PACKED") == nullptr;
    if (scalar_pair_tiles)
        k_pair_force<POT><<<blocks, 32*WPB, 0, stream>>>(ctx->n, ctx->g_lo, ctx->g_hi, ctx->xq, ctx->par[pf.set],
                                                              L.entries, L.counts, L.gflags, L.cap, out,
                                                              accumulate ? 1 : 0, pot, rc2, bb, boxf);
    else
        k_pair_force2<typename PackedOf<POT>::type><<<blocks, 32*WPB, 0, stream>>>(
            ctx->n, ctx->g_lo, ctx->g_hi, ctx->xq, ctx->par[pf.set], L.entries, L.counts, L.gflags, L.cap, out,
            accumulate ? 1 : 0, typename PackedOf<POT>::type{pot.p}, rc2, bb, boxf);
    if (ctx->profiling) {
        cudaEventRecord(ev1, stream);
        ctx->prof_events.push_back(ev0); ctx->prof_events.push_back(ev1);
        ctx->prof_tags.push_back((int)(&pf - ctx->pair_forces.data()));
    }
    ctx->counters[2]++;
    B2_LAUNCH_CHECK();
    // ~2e-3 band pairs per atom: one thread each, at least 8 blocks
    const int owned = ctx->a_hi - ctx->a_lo;
    const int band_blocks = std::min(296, std::max(8, owned/16384));
    const bool fused_apply = owned < 400000;       // ~800 band pairs: the serial tail is shorter than a launch
    k_pair_band<POTD><<<band_blocks, 128, 0, stream>>>(ctx->x, ctx->pard[pf.set], bb, potd, rcd*rcd, ctx->box[0],
                                                      ctx->box[1], ctx->box[2], out, fused_apply ? 1 : 0);
    B2_LAUNCH_CHECK();
    if (!fused_apply) {
        k_pair_band_apply<<<band_blocks, 128, 0, stream>>>(bb, out);
        B2_LAUNCH_CHECK();
    }
    return B2_OK;
}

template <class POT, bool SOFT>
static int launch_energy(b2_context* ctx, const PairForce& pf, POT pot, double rc2, double* acc) {
    const NList& L = ctx->lists[pf.list];
    const int blocks = (ctx->g_hi - ctx->g_lo + WPB - 1)/WPB;
    if (blocks == 0) return B2_OK;
    k_pair_energy<POT, SOFT><<<blocks, 32*WPB, 0, ctx->stream>>>(ctx->n, ctx->g_lo, ctx->g_hi, ctx->a_lo, ctx->a_hi, ctx->x,
                                                                  ctx->pard[pf.set], L.entries, L.counts,
                                                                  L.cap, pot, rc2, ctx->box[0], ctx->box[1],
                                                                  ctx->box[2], acc, ctx->xq);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

#define DISPATCH(T, CALL_LJC, CALL_SOFT)                                                                    \
    switch (pf.family) {                                                                                    \
    case B2_PAIR_NEAR: {                                                                                    \
        const int v = (int)pf.params[0];                                                                    \
        if (v == 0) { LJCPot<COUL_PLAIN, LJ_STD, SW_ALL, SWF_LINEAR, VAR_NONE, T> pot{p}; CALL_LJC; }        \
        else if (v == 1) { LJCPot<COUL_PLAIN, LJ_STD, SW_ALL, SWF_LINEAR, VAR_SHIFT, T> pot{p}; CALL_LJC; }  \
        else { LJCPot<COUL_PLAIN, LJ_STD, SW_ALL, SWF_LINEAR, VAR_FSWITCH, T> pot{p}; CALL_LJC; }            \
        break;                                                                                              \
    }                                                                                                       \
    case B2_PAIR_DAMPED:                                                                                    \
        if ((int)pf.params[3] == 1) { LJCPot<COUL_ERFC, LJ_STD, SW_ALL, SWF_LINEAR, VAR_NONE, T> pot{p}; CALL_LJC; } \
        else { LJCPot<COUL_ERFC, LJ_STD, SW_ALL, SWF_POWER, VAR_NONE, T> pot{p}; CALL_LJC; }                 \
        break;                                                                                              \
    case B2_PAIR_LJC: {                                                                                     \
        const int ck = (int)pf.params[1];                                                                   \
        const bool sw = pf.params[5] != 0.0;       /* OpenMM switch on the LJ part in use? */               \
        if (ck == 2 && !sw) { LJCPot<COUL_RF, LJ_STD, SW_NONE, SWF_LINEAR, VAR_NONE, T> pot{p}; CALL_LJC; }  \
        else if (ck == 3 && !sw) { LJCPot<COUL_ERFC, LJ_STD, SW_NONE, SWF_LINEAR, VAR_NONE, T> pot{p}; CALL_LJC; } \
        else if (ck == 2) { LJCPot<COUL_RF, LJ_STD, SW_LJ, SWF_LINEAR, VAR_NONE, T> pot{p}; CALL_LJC; }      \
        else if (ck == 3) { LJCPot<COUL_ERFC, LJ_STD, SW_LJ, SWF_LINEAR, VAR_NONE, T> pot{p}; CALL_LJC; }    \
        /* ck == 1 is a CustomNonbondedForce: OpenMM's switch multiplies its WHOLE energy, Coulomb included */ \
        else if (sw) { LJCPot<COUL_PLAIN, LJ_STD, SW_ALL, SWF_LINEAR, VAR_NONE, T> pot{p}; CALL_LJC; }      \
        else { LJCPot<COUL_PLAIN, LJ_STD, SW_NONE, SWF_LINEAR, VAR_NONE, T> pot{p}; CALL_LJC; }             \
        break;                                                                                              \
    }                                                                                                       \
    case B2_PAIR_LJ_VIRIAL: { LJCPot<COUL_NONE, LJ_VIRIAL, SW_ALL, SWF_LINEAR, VAR_NONE, T> pot{p}; CALL_LJC; break; } \
    case B2_PAIR_SOFTCORE: { SoftcorePot<T> pot{p}; CALL_SOFT; break; }                                      \
    default: return b2_fail(ctx, B2_ERR_UNSUPPORTED, "unknown pair family %d", pf.family);                  \
    }

template <class POT>
struct DoubleOf;
template <int A, int B, int C, int D, int E>
struct DoubleOf<LJCPot<A, B, C, D, E, float>> { typedef LJCPot<A, B, C, D, E, double> type; };
template <>
struct DoubleOf<SoftcorePot<float>> { typedef SoftcorePot<double> type; };

int pair_eval_forces(b2_context* ctx, const PairForce& pf, float4* out, bool accumulate, int lane) {
    float rc2, unused;
    PotParams<float> p = make_params<float>(ctx, pf, &rc2);
    PotParams<double> pd = make_params<double>(ctx, pf, &unused);
#define CALL_FORCE { typename DoubleOf<decltype(pot)>::type potd{pd}; B2_TRY(launch_force(ctx, pf, pot, potd, rc2, out, accumulate, lane)); }
    DISPATCH(float, CALL_FORCE, CALL_FORCE);
#undef CALL_FORCE
    return B2_OK;
}

int pair_eval_energy(b2_context* ctx, const PairForce& pf, int group) {
    float rc2f;
    PotParams<double> p = make_params<double>(ctx, pf, &rc2f);
    double rc = pf.cutoff;
    if (pf.family == B2_PAIR_NEAR) rc = std::min(rc, pf.params[2]);
    if (pf.family == B2_PAIR_DAMPED) rc = std::min(rc, pf.params[2]);
    const double rc2 = rc*rc;
    // accumulator block: [0] e, [1] w, [2] dlv, [3] dlc  (scratch at d_energy+72), folded by caller
    double* acc = ctx->d_energy + 72;
    B2_CUDA(cudaMemsetAsync(acc, 0, 4*sizeof(double), ctx->stream));
    DISPATCH(double, B2_TRY((launch_energy<decltype(pot), false>(ctx, pf, pot, rc2, acc))),
             B2_TRY((launch_energy<decltype(pot), true>(ctx, pf, pot, rc2, acc))));
    (void)group;
    return B2_OK;
}

template <class POT>
static int launch_force64(b2_context* ctx, const PairForce& pf, POT pot, double rc2, double* out) {
    const NList& L = ctx->lists[pf.list];
    const int blocks = (ctx->g_hi - ctx->g_lo + WPB - 1)/WPB;
    if (blocks == 0) return B2_OK;
    k_pair_force64<POT><<<blocks, 32*WPB, 0, ctx->stream>>>(ctx->n, ctx->g_lo, ctx->g_hi, ctx->x, ctx->pard[pf.set], L.entries,
                                                          L.counts, L.cap, pot, rc2, ctx->box[0], ctx->box[1], ctx->box[2],
                                                          out);
    B2_LAUNCH_CHECK();
    return B2_OK;
}

// float64 forces of one pair force accumulated into out[n][3] (engine order)
int pair_eval_forces64(b2_context* ctx, const PairForce& pf, double* out) {
    float rc2f;
    PotParams<double> p = make_params<double>(ctx, pf, &rc2f);
    const double rc = effective_cutoff(pf);
    DISPATCH(double, B2_TRY(launch_force64(ctx, pf, pot, rc*rc, out)), B2_TRY(launch_force64(ctx, pf, pot, rc*rc, out)));
    return B2_OK;
}

int pair_count_set(b2_context* ctx, const PairForce& pf, long long* count, unsigned long long* checksum,
                   int* pairs_dev, long long capacity) {
    if (ctx->nranks > 1) return b2_fail(ctx, B2_ERR_UNSUPPORTED, "pair-set extraction is a single-GPU diagnostic");
    const NList& L = ctx->lists[pf.list];
    double rc = pf.cutoff;
    if (pf.family == B2_PAIR_NEAR || pf.family == B2_PAIR_DAMPED) rc = std::min(rc, pf.params[2]);
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(ctx->d_energy + 76);
    B2_CUDA(cudaMemsetAsync(acc, 0, 3*sizeof(unsigned long long), ctx->stream));
    const int blocks = (ctx->ngroups + WPB - 1)/WPB;
    k_pair_set<<<blocks, 32*WPB, 0, ctx->stream>>>(ctx->n, ctx->ngroups, ctx->x, ctx->orig, L.entries, L.counts,
                                                    L.cap, rc*rc, ctx->box[0], ctx->box[1], ctx->box[2], acc,
                                                    pairs_dev, capacity);
    B2_LAUNCH_CHECK();
    unsigned long long h[3];
    B2_CUDA(cudaMemcpyAsync(h, acc, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    *count = (long long)h[0];
    *checksum = h[1];
    return B2_OK;
}

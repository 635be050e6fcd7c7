// Per-term device functions of the explicit-list (bonded) families, shared by the stand-alone
// bonded kernels (bonded.cu) and the fused RESPA inner-loop kernel (integrate.cu).
//
// GEO abstracts where positions come from and where forces go:
//   GlobalGeo  master positions in HBM (sorted order), fp32 atomics into a force buffer
//   LocalGeo   a molecule chunk staged in shared memory, fp64 atomics into shared memory
#pragma once

#include <math.h>

#include "ctx.h"
#include "vm.cuh"

struct BondArgs {
    int nterms, stride, periodic, family;
    const int* atoms;
    const double* params;
    const int* inv;
    int a_lo, a_hi;              // owned atoms: a term belongs to the owner of its first atom
    double box[3];
    double g[8];
    const int* code_e; int ncode_e;
    const int* code_de; int ncode_de;
    const double* consts;
};

// Forces go through a 64-bit fixed-point accumulator per degree of freedom (2^-32 kJ/mol/nm): integer
// addition is associative, so the result does not depend on the order in which the terms of an atom arrive
// (float atomics would make trajectories irreproducible); k_fixed_to_force adds the sums to the fp32 buffer.
#define B2_FIXED_SCALE 4294967296.0
struct GlobalGeo {
    const double* x;
    unsigned long long* acc;     // [n][3]
    __device__ __forceinline__ double pos(int i, int k) const { return x[3*i+k]; }
    __device__ __forceinline__ void add(int i, double fx, double fy, double fz) const {
        atomicAdd(&acc[3*(size_t)i], (unsigned long long)__double2ll_rn(fx*B2_FIXED_SCALE));
        atomicAdd(&acc[3*(size_t)i+1], (unsigned long long)__double2ll_rn(fy*B2_FIXED_SCALE));
        atomicAdd(&acc[3*(size_t)i+2], (unsigned long long)__double2ll_rn(fz*B2_FIXED_SCALE));
    }
};

// float64 force buffer [n][3] (report cadence)
struct GlobalGeo64 {
    const double* x;
    double* out;
    __device__ __forceinline__ double pos(int i, int k) const { return x[3*i+k]; }
    __device__ __forceinline__ void add(int i, double fx, double fy, double fz) const {
        atomicAdd(&out[3*i], fx);
        atomicAdd(&out[3*i+1], fy);
        atomicAdd(&out[3*i+2], fz);
    }
};

// the same fixed-point accumulation in shared memory (fused inner loop)
struct LocalGeo {
    const double* xs;            // [chunk][3] shared memory
    unsigned long long* fs;      // [chunk][3] shared memory, fixed point
    int base;                    // sorted index of the chunk's first atom
    __device__ __forceinline__ double pos(int i, int k) const { return xs[3*(i - base)+k]; }
    // 64-bit integer atomics on shared memory compile to compare-and-swap loops; two native 32-bit adds with the
    // carry of the low word forwarded to the high word give the same sum (every add contributes its carry exactly
    // once, in whatever order the adds arrive) -- still associative, still order-independent
    __device__ __forceinline__ void add1(unsigned long long* cell, double f) const {
        const unsigned long long q = (unsigned long long)__double2ll_rn(f*B2_FIXED_SCALE);
        const unsigned lo = (unsigned)q, hi = (unsigned)(q >> 32);
        unsigned* w = reinterpret_cast<unsigned*>(cell);         // little endian: w[0] low word, w[1] high word
        const unsigned old = atomicAdd(w, lo);
        const unsigned carry = (old + lo) < old ? 1u : 0u;
        if (hi + carry) atomicAdd(w + 1, hi + carry);
    }
    __device__ __forceinline__ void add(int i, double fx, double fy, double fz) const {
        add1(&fs[3*(i - base)], fx);
        add1(&fs[3*(i - base)+1], fy);
        add1(&fs[3*(i - base)+2], fz);
    }
    __device__ __forceinline__ double force(int local, int k) const {
        return (double)(long long)fs[3*local+k]*(1.0/B2_FIXED_SCALE);
    }
};

template <class GEO>
__device__ __forceinline__ void delta(const BondArgs& a, const GEO& geo, int i, int j, double (&d)[3]) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
        d[k] = geo.pos(j, k) - geo.pos(i, k);
        if (a.periodic) d[k] -= a.box[k]*rint(d[k]/a.box[k]);
    }
}

__device__ __forceinline__ void block_accumulate(double e, double w, double* acc) {
    for (int o = 16; o > 0; o >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, o);
        w += __shfl_xor_sync(0xffffffffu, w, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (e != 0.0) atomicAdd(&acc[0], e);
        if (w != 0.0) atomicAdd(&acc[1], w);
    }
}

// two-body families: E(r); force from dE/dr
// CUSTOM = false compiles the bytecode interpreter of CustomBondForce / CustomAngleForce terms out (the fused
// inner loop of systems without such terms: no evaluation stack in local memory, fewer registers)
// Every term function comes in two parts: term_x() resolves the term (atom indices in the engine's order through
// two dependent table look-ups, parameter pointer) and term_x_core() evaluates it.  The fused inner loop resolves
// a thread's term ONCE per launch and evaluates it at every inner iteration from registers.
template <bool FORCE, bool ENERGY, bool CUSTOM = true, class GEO>
__device__ __forceinline__ void term_bond2_core(const BondArgs& a, int i, int j, const double* p, const GEO& geo,
                                                double& e, double& w) {
    {
        double d[3];
        delta(a, geo, i, j, d);
        const double r2 = d[0]*d[0] + d[1]*d[1] + d[2]*d[2];
        const double ir = rsqrt(r2);         // one reciprocal square root instead of a square root and divisions
        const double r = r2*ir;
        double dedr = 0;
        if (a.family == B2_BOND_HARMONIC) {
            const double dr = r - p[0];
            e = 0.5*p[1]*dr*dr;
            dedr = p[1]*dr;
        } else if (a.family == B2_BOND_LJC) {
            const double qq = p[0], sig = p[1], eps = p[2];
            const double s2 = sig*sig*ir*ir, s6 = s2*s2*s2;
            e = 4*eps*s6*(s6 - 1) + a.g[0]*qq*ir;
            dedr = -(24*eps*s6*(2*s6 - 1) + a.g[0]*qq*ir)*ir;
            if (a.g[1] > 0) {
                const double al = a.g[1], kq = a.g[0]*p[3];
                const double er = erf(al*r);
                e -= kq*er*ir;
                dedr -= kq*(2*al/sqrt(M_PI)*exp(-al*al*r2)*ir - er*ir*ir);
            }
        } else if constexpr (CUSTOM) {
            double vars[10];
            vars[0] = r;
            for (int k = 0; k < a.stride && k < 9; k++) vars[1+k] = p[k];
            if (ENERGY) e = vm_run<2>(a.code_e, a.ncode_e, a.consts, nullptr, nullptr, 0, vars, nullptr, nullptr);
            dedr = vm_run<2>(a.code_de, a.ncode_de, a.consts, nullptr, nullptr, 0, vars, nullptr, nullptr);
        }
        w = -dedr*r;
        if (FORCE) {
            const double s = dedr*ir;   // force on j = -dE/dr * d/r
            geo.add(i, s*d[0], s*d[1], s*d[2]);
            geo.add(j, -s*d[0], -s*d[1], -s*d[2]);
        }
    }
}

template <bool FORCE, bool ENERGY, bool CUSTOM = true, class GEO>
__device__ __forceinline__ void term_bond2(const BondArgs& a, int t, const GEO& geo, double& e, double& w) {
    const int i = a.inv[a.atoms[2*t]], j = a.inv[a.atoms[2*t+1]];
    if (i < a.a_lo || i >= a.a_hi) return;
    term_bond2_core<FORCE, ENERGY, CUSTOM>(a, i, j, a.params + (size_t)t*a.stride, geo, e, w);
}

template <bool FORCE, bool ENERGY, bool CUSTOM = true, class GEO>
__device__ __forceinline__ void term_angle_core(const BondArgs& a, int i, int j, int k, const double* p, const GEO& geo,
                                                double& e) {
    {
        double u[3], v[3];
        delta(a, geo, j, i, u);
        delta(a, geo, j, k, v);
        const double ru2 = u[0]*u[0] + u[1]*u[1] + u[2]*u[2], rv2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
        const double iu = rsqrt(ru2), iv = rsqrt(rv2);       // division-free: three reciprocal square roots per angle
        const double iuv = iu*iv;
        double c = (u[0]*v[0] + u[1]*v[1] + u[2]*v[2])*iuv;
        c = fmin(1.0, fmax(-1.0, c));
        const double theta = acos(c);
        double dedt = 0;
        if (a.family == B2_ANGLE_HARMONIC) {
            const double dt = theta - p[0];
            e = 0.5*p[1]*dt*dt;
            dedt = p[1]*dt;
        } else if constexpr (CUSTOM) {
            double vars[10];
            vars[0] = theta;
            for (int q = 0; q < a.stride && q < 9; q++) vars[1+q] = p[q];
            if (ENERGY) e = vm_run<2>(a.code_e, a.ncode_e, a.consts, nullptr, nullptr, 0, vars, nullptr, nullptr);
            dedt = vm_run<2>(a.code_de, a.ncode_de, a.consts, nullptr, nullptr, 0, vars, nullptr, nullptr);
        }
        if (FORCE) {
            const double is = rsqrt(fmax(1.0 - c*c, 1e-30));
            // d theta/d r_i = -(v/(ru rv) - c u/ru^2)/sin(theta);  force = -dE/dtheta * d theta/d r
            const double g = dedt*is, ciu2 = c*iu*iu, civ2 = c*iv*iv;
            double fi[3], fk[3];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                fi[q] = g*(v[q]*iuv - ciu2*u[q]);
                fk[q] = g*(u[q]*iuv - civ2*v[q]);
            }
            geo.add(i, fi[0], fi[1], fi[2]);
            geo.add(k, fk[0], fk[1], fk[2]);
            geo.add(j, -(fi[0]+fk[0]), -(fi[1]+fk[1]), -(fi[2]+fk[2]));
        }
    }
}

template <bool FORCE, bool ENERGY, bool CUSTOM = true, class GEO>
__device__ __forceinline__ void term_angle(const BondArgs& a, int t, const GEO& geo, double& e) {
    const int i = a.inv[a.atoms[3*t]], j = a.inv[a.atoms[3*t+1]], k = a.inv[a.atoms[3*t+2]];
    if (i < a.a_lo || i >= a.a_hi) return;
    term_angle_core<FORCE, ENERGY, CUSTOM>(a, i, j, k, a.params + (size_t)t*a.stride, geo, e);
}

__device__ __forceinline__ void cross3(const double (&a)[3], const double (&b)[3], double (&c)[3]) {
    c[0] = a[1]*b[2] - a[2]*b[1];
    c[1] = a[2]*b[0] - a[0]*b[2];
    c[2] = a[0]*b[1] - a[1]*b[0];
}

template <bool FORCE, bool ENERGY, class GEO>
__device__ __forceinline__ void term_torsion_core(const BondArgs& a, int a1, int a2, int a3, int a4, const double* p,
                                                  const GEO& geo, double& e) {
    {
        double F[3], G[3], H[3], A[3], B[3], C[3];
        delta(a, geo, a2, a1, F);   // r1 - r2
        delta(a, geo, a3, a2, G);   // r2 - r3
        delta(a, geo, a3, a4, H);   // r4 - r3
        cross3(F, G, A);
        cross3(H, G, B);
        cross3(B, A, C);
        const double A2 = A[0]*A[0] + A[1]*A[1] + A[2]*A[2], B2 = B[0]*B[0] + B[1]*B[1] + B[2]*B[2];
        const double G2 = G[0]*G[0] + G[1]*G[1] + G[2]*G[2], gn = sqrt(G2);
        const double norm = sqrt(A2*B2);
        const double cosphi = (A[0]*B[0] + A[1]*B[1] + A[2]*B[2])/norm;
        const double sinphi = (C[0]*G[0] + C[1]*G[1] + C[2]*G[2])/(norm*gn);
        const double phi = atan2(sinphi, cosphi);
        const double n = p[0], phase = p[1], k = p[2];
        e = k*(1.0 + cos(n*phi - phase));
        if (FORCE) {
            const double dedphi = -k*n*sin(n*phi - phase);
            const double fg = F[0]*G[0] + F[1]*G[1] + F[2]*G[2], hg = H[0]*G[0] + H[1]*G[1] + H[2]*G[2];
            double f1[3], f2[3], f3[3], f4[3];
            for (int q = 0; q < 3; q++) {
                const double d1 = -gn/A2*A[q];
                const double d4 = gn/B2*B[q];
                const double d2 = gn/A2*A[q] + fg/(A2*gn)*A[q] - hg/(B2*gn)*B[q];
                const double d3 = -gn/B2*B[q] - fg/(A2*gn)*A[q] + hg/(B2*gn)*B[q];
                f1[q] = -dedphi*d1; f2[q] = -dedphi*d2; f3[q] = -dedphi*d3; f4[q] = -dedphi*d4;
            }
            geo.add(a1, f1[0], f1[1], f1[2]);
            geo.add(a2, f2[0], f2[1], f2[2]);
            geo.add(a3, f3[0], f3[1], f3[2]);
            geo.add(a4, f4[0], f4[1], f4[2]);
        }
    }
}

template <bool FORCE, bool ENERGY, class GEO>
__device__ __forceinline__ void term_torsion(const BondArgs& a, int t, const GEO& geo, double& e) {
    const int a1 = a.inv[a.atoms[4*t]], a2 = a.inv[a.atoms[4*t+1]], a3 = a.inv[a.atoms[4*t+2]], a4 = a.inv[a.atoms[4*t+3]];
    if (a1 < a.a_lo || a1 >= a.a_hi) return;
    term_torsion_core<FORCE, ENERGY>(a, a1, a2, a3, a4, a.params + (size_t)t*a.stride, geo, e);
}


// all explicit-list forces of one force group in a single launch (force path only)
#define B2_MAX_BATCH 8
struct BondBatch {
    int count;
    int first[B2_MAX_BATCH + 1];     // prefix sums of term counts
    int arity[B2_MAX_BATCH];
    BondArgs a[B2_MAX_BATCH];
};


BondArgs bonded_make_args(b2_context* ctx, const BondedForce& bf);

// Closed-form pair potentials (kernel K3 families of SURVEY 2.1), templated on the scalar type:
// float for the force kernels on the hot path, double for energies / virials at report cadence.
//
// Every functor returns, for one pair at squared distance r2 (already inside the cutoff):
//     rF = -r dE/dr     (so the force on atom i is  rF / r^2 * (x_i - x_j)  and the pair virial is rF)
//     e  = E(r)         (only meaningful in the double instantiation for the force-switch family)
// with Lorentz-Berthelot mixing done by the caller: sig = (sigma_i+sigma_j)/2,
// eps = sqrt(eps_i eps_j), qq = q_i q_j (forces.py:247-258).
//
// Energy strings restated (reference file:line):
//   near none / shift / force-switch           forces.py:541-563
//   damped-smoothed, degree 1 and d >= 2       forces.py:448-455
//   LJ + {plain, reaction-field, erfc} Coulomb openmm.NonbondedForce as set up in forces.py:152-190
//   LJ virial 24 eps (2 s^12 - s^6)            systems.py:894
//   soft-core LJ + scaled Coulomb              forces.py:749-750,785-786
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#include "packed.cuh"

enum { COUL_NONE = 0, COUL_PLAIN = 1, COUL_RF = 2, COUL_ERFC = 3 };
enum { LJ_STD = 0, LJ_VIRIAL = 1 };
enum { SW_NONE = 0, SW_ALL = 1, SW_LJ = 2 };
enum { SWF_LINEAR = 0, SWF_POWER = 1 };
enum { VAR_NONE = 0, VAR_SHIFT = 1, VAR_FSWITCH = 2 };

B2_HD float b2_rsqrt(float x) {
#ifdef __CUDA_ARCH__
    return rsqrtf(x);
#else
    return 1.0f/sqrtf(x);
#endif
}
B2_HD double b2_rsqrt(double x) { return 1.0/sqrt(x); }
// 1/r and 1/r^2 from r^2.  fp32: MUFU.RSQ (<= 2 ulp) refined by one Newton step (~0.5 ulp), because the
// r^-12 term amplifies the relative error of 1/r twelve-fold.
B2_HD void b2_inverse(float r2, float& rinv, float& rinv2) {
    float y;
#ifdef __CUDA_ARCH__
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(r2));   // r2 is never subnormal: no range fix-up code
#else
    y = 1.0f/sqrtf(r2);
#endif
    const float e = fmaf(-r2*y, y, 1.0f);
    y = fmaf(0.5f*y, e, y);
    rinv = y;
    rinv2 = y*y;
}
B2_HD void b2_inverse(double r2, double& rinv, double& rinv2) {
#ifdef __CUDA_ARCH__
    rinv = rsqrt(r2);              // one reciprocal square root (<= 1 ulp) instead of a square root and two divisions
    rinv2 = rinv*rinv;
#else
    rinv = 1.0/sqrt(r2);
    rinv2 = 1.0/r2;
#endif
}
B2_HD float b2_erfc(float x) { return erfcf(x); }
B2_HD double b2_erfc(double x) { return erfc(x); }
B2_HD float b2_exp(float x) {
#ifdef __CUDA_ARCH__
    return __expf(x);
#else
    return expf(x);
#endif
}
B2_HD double b2_exp(double x) { return exp(x); }
B2_HD float b2_log(float x) {
#ifdef __CUDA_ARCH__
    return __logf(x);
#else
    return logf(x);
#endif
}
B2_HD double b2_log(double x) { return log(x); }
B2_HD float b2_max(float a, float b) { return fmaxf(a, b); }
B2_HD double b2_max(double a, double b) { return fmax(a, b); }

template <typename T>
struct PotParams {
    T kc;            // Coulomb constant times sign handled by caller? no: plain Kc
    T sign;          // +1 / -1 (subtract=True, discount forces)
    T rs, iw;        // switch start and 1/(rc-rs)  (iw = 0 disables the switch)
    T rsd, iwd;      // rs^d and 1/(rc^d - rs^d) for the power-law switching variable
    int degree;
    T krf, crf;      // reaction field
    T alpha, tasp;   // Ewald/damping alpha and 2 alpha/sqrt(pi)
    T inv_rc0;       // 1/rc0 (shift variant)
    T b, f12c, f6c, f1c, c12, c6, c1;   // force-switch energy constants (forces.py:552-563)
    T lam_v, lam_c;  // soft-core
    T gmode;         // soft-core: 1 = interaction-group mode, charges are +-1 set labels and only pairs of
                     // unlike labels interact (OpenMM addInteractionGroup(A, complement of A))
    const double* lam_v_dev;   // when not null the couplings are read from device memory at run time:
    const double* lam_c_dev;   // integrator globals (AFED extended variables), no host round trip
};

template <typename T>
B2_HD void switch_eval(const PotParams<T>& p, int swf, T r, T r2, T& S, T& rdS) {
    T u, rdu;
    if (swf == SWF_LINEAR) {
        u = b2_max((r - p.rs)*p.iw, T(0));
        rdu = r*p.iw;
    } else {
        T rd = r2;                                  // r^d by repeated multiplication, d >= 2
        for (int k = 2; k < p.degree; k++) rd *= r;
        u = b2_max((rd - p.rsd)*p.iwd, T(0));
        rdu = T(p.degree)*rd*p.iwd;
    }
    const T u2 = u*u;
    S = T(1) - u2*u*(T(10) - T(15)*u + T(6)*u2);
    const T om = T(1) - u;
    rdS = T(-30)*u2*om*om*rdu;                       // r dS/dr
}

template <int COUL, int LJ, int SW, int SWF, int VAR, typename T>
struct LJCPot {
    PotParams<T> p;

    template <bool WANT_E>
    B2_HD void operator()(T r2, T qq, T sig, T eps, T& rF, T& e, T& rinv2) const {
        T rinv;
        b2_inverse(r2, rinv, rinv2);
        const T r = r2*rinv;
        const T s2 = sig*sig*rinv2;
        const T s6 = s2*s2*s2;
        T elj, rflj;
        if (LJ == LJ_STD) {
            elj = T(4)*eps*s6*(s6 - T(1));
            rflj = T(24)*eps*s6*(T(2)*s6 - T(1));
        } else {
            elj = T(24)*eps*s6*(T(2)*s6 - T(1));
            rflj = T(144)*eps*s6*(T(4)*s6 - T(1));
        }
        T ec = T(0), rfc = T(0);
        const T kqq = p.kc*qq;
        if (COUL == COUL_PLAIN) {
            ec = kqq*rinv;
            rfc = ec;
        } else if (COUL == COUL_RF) {
            ec = kqq*(rinv + p.krf*r2 - p.crf);
            rfc = kqq*(rinv - T(2)*p.krf*r2);
        } else if (COUL == COUL_ERFC) {
            const T ar = p.alpha*r;
            const T erfc_r = b2_erfc(ar)*rinv;
            ec = kqq*erfc_r;
            rfc = kqq*(erfc_r + p.tasp*b2_exp(-ar*ar));
        }
        if (SW == SW_NONE) {
            rF = rflj + rfc;
            e = elj + ec;
        } else if (SW == SW_LJ) {
            T S, rdS;
            switch_eval(p, SWF, r, r2, S, rdS);
            rF = S*rflj - rdS*elj + rfc;
            e = S*elj + ec;
        } else {
            T S, rdS;
            switch_eval(p, SWF, r, r2, S, rdS);
            if (VAR == VAR_NONE) {
                const T V = elj + ec;
                rF = S*(rflj + rfc) - rdS*V;
                e = S*V;
            } else if (VAR == VAR_SHIFT) {
                const T c2 = sig*sig*p.inv_rc0*p.inv_rc0;
                const T c6 = c2*c2*c2;
                const T V = elj + ec - (T(4)*eps*c6*(c6 - T(1)) + kqq*p.inv_rc0);
                rF = S*(rflj + rfc) - rdS*V;
                e = S*V;
            } else {
                rF = S*(rflj + rfc);
                e = T(0);
                if (WANT_E) {
                    // V*(r) - V*(rc0): f_n(u) from forces.py:552-554 (well conditioned only in double)
                    const T u = b2_max((r - p.rs)*p.iw, T(0));
                    T f12 = T(1), f6 = T(1), f1 = T(1);
                    if (u > T(0)) {
                        const T b = p.b, R = u/b + T(1);
                        const T u2 = u*u, u3 = u2*u, u4 = u2*u2, u5 = u4*u, b2 = b*b, b3 = b2*b;
                        const T R2 = R*R, R4 = R2*R2, R6 = R4*R2, R12 = R6*R6;
                        f12 += (T(6)*b2 - T(21)*b + T(28))*(b3*(R12 - T(1)) - T(12)*b2*u - T(66)*b*u2 - T(220)*u3)/T(462)
                               + T(45)*(T(7) - T(2)*b)*u4/T(14) - T(72)*u5/T(7);
                        f6 += (T(6)*b2 - T(3)*b + T(1))*(b3*(R6 - T(1)) - T(6)*b2*u - T(15)*b*u2 - T(20)*u3)
                              + T(45)*(T(1) - T(2)*b)*u4 - T(36)*u5;
                        f1 += T(5)*(b + T(1))*(b + T(1))*(T(6)*b3*R*b2_log(R) - T(6)*b2*u - T(3)*b*u2 + u3)
                              + u4*(T(3)*u - T(5)*b - T(10))/T(2);
                    }
                    const T c2 = sig*sig*p.inv_rc0*p.inv_rc0;
                    const T c6 = c2*c2*c2;
                    e = T(4)*eps*(f12*s6*s6 - f6*s6) + kqq*f1*rinv
                        - (T(4)*eps*(p.f12c*c6*c6 - p.f6c*c6) + kqq*p.f1c*p.inv_rc0);
                }
            }
        }
        rF *= p.sign;
        e *= p.sign;
    }
};

// Beutler soft core:  E = S(r) [ 4 lam_v eps (1-x)/x^2 + Kc lam_c qq / r ],  x = (r/sig)^6 + (1-lam_v)/2
template <typename T>
struct SoftcorePot {
    PotParams<T> p;

    template <bool WANT_E>
    B2_HD void operator()(T r2, T qq, T sig, T eps, T& rF, T& e, T& rinv2) const {
        T dv, dc;
        eval(r2, qq, sig, eps, rF, e, dv, dc);
        rinv2 = T(1)/r2;
    }

    B2_HD void eval(T r2, T qq, T sig, T eps, T& rF, T& e, T& dEdlv, T& dEdlc) const {
        const T lam_v = p.lam_v_dev ? T(*p.lam_v_dev) : p.lam_v;
        const T lam_c = p.lam_c_dev ? T(*p.lam_c_dev) : p.lam_c;
        T pair_scale = T(1);
        if (p.gmode != T(0)) {            // qq = +1 (same set: no interaction) or -1 (solute-solvent)
            pair_scale = T(0.5)*(T(1) - qq);
            qq = T(0);
        }
        const T rinv = b2_rsqrt(r2);
        const T r = r2*rinv;
        const T is2 = T(1)/(sig*sig);
        const T q2 = r2*is2;
        const T r6s = q2*q2*q2;
        const T x = r6s + T(0.5)*(T(1) - lam_v);
        const T ix = T(1)/x;
        const T g = (T(1) - x)*ix*ix;              // (1-x)/x^2
        const T dg = (x - T(2))*ix*ix*ix;          // d/dx
        const T elj = T(4)*lam_v*eps*g;
        const T rflj = -T(4)*lam_v*eps*dg*T(6)*r6s;
        const T ec = p.kc*lam_c*qq*rinv;
        T S = T(1), rdS = T(0);
        if (p.iw != T(0)) switch_eval(p, SWF_LINEAR, r, r2, S, rdS);
        const T V = elj + ec;
        rF = S*(rflj + ec) - rdS*V;
        e = S*V;
        dEdlv = S*(T(4)*eps*g - T(2)*lam_v*eps*dg);
        dEdlc = S*p.kc*qq*rinv;
        rF *= pair_scale; e *= pair_scale; dEdlv *= pair_scale; dEdlc *= pair_scale;
    }
};

// ---------------------------------------------------------------------------------------------
// Packed fp32x2 force functor (the B2_PAIR_PACKED=1 tiles; measured slower than the scalar tiles on B200, see
// pair.cu).  sm_100a has FFMA2 / FADD2 / FMUL2: one instruction issue does the arithmetic of TWO pairs.  Every
// lane of a packed tile evaluates two list slots at once; the closed forms above are restated on a two-wide value
// type with explicit fused multiply-adds (the *_rn intrinsics are never contracted by the compiler).
// Only what the force kernels need (no energies) -- energies / virials stay in the float64 functors.
// Compiles for the host too (component-wise fmaf), so the restatement is checked against the
// scalar functors without a GPU (tests/test_packed_potentials.py).
// ---------------------------------------------------------------------------------------------

template <int COUL, int LJ, int SW, int SWF, int VAR>
struct LJCForce2 {
    PotParams<float> p;
    // the caller folds these into the i-atom's parameters once per tile
    static constexpr float EPS_SCALE = (LJ == LJ_STD) ? 24.f : 144.f;
    B2_HD float charge_scale() const { return p.kc; }      // Kc is folded into q_i as well

    // r2: squared distances of two pairs; kqq = Kc q_i q_j; sig = (sigma_i + sigma_j)/2;
    // eps = EPS_SCALE sqrt(eps_i eps_j).  Returns  -(dE/dr)/r  (times sign): force on i = result * (x_i - x_j).
    B2_HD F2 operator()(F2 r2, F2 kqq, F2 sig, F2 eps) const {
        // 1/r: MUFU.RSQ per half, one packed Newton step (the r^-12 term amplifies the error twelve-fold)
        F2 y = f2(b2_rsqrt_approx(r2.v.x), b2_rsqrt_approx(r2.v.y));
        const F2 t = r2*y;
        const F2 e = fma2(t, y, f2(-1.f));             // r2 y^2 - 1
        y = fma2(y*f2(-0.5f), e, y);
        const F2 rinv = y, rinv2 = y*y;
        const F2 sr = sig*y;
        const F2 s2 = sr*sr;
        const F2 s6 = s2*s2*s2;
        const F2 es6 = eps*s6;
        F2 rflj, elj;
        if (LJ == LJ_STD) {
            rflj = es6*fma2(s6, f2(2.f), f2(-1.f));                     // 24 eps s6 (2 s6 - 1)
            elj = (es6*f2(1.f/6.f))*(s6 + f2(-1.f));                    // 4 eps s6 (s6 - 1)
        } else {
            rflj = es6*fma2(s6, f2(4.f), f2(-1.f));                     // 144 eps s6 (4 s6 - 1)
            elj = (es6*f2(1.f/6.f))*fma2(s6, f2(2.f), f2(-1.f));        // 24 eps s6 (2 s6 - 1)
        }
        F2 ec = f2(0.f), rfc = f2(0.f);
        if (COUL == COUL_PLAIN) {
            ec = kqq*rinv;
            rfc = ec;
        } else if (COUL == COUL_RF) {
            ec = kqq*(fma2(r2, f2(p.krf), rinv) + f2(-p.crf));
            rfc = kqq*fma2(r2, f2(-2.f*p.krf), rinv);
        } else if (COUL == COUL_ERFC) {
            const F2 r = r2*rinv;
            const F2 ar = r*f2(p.alpha);
            const F2 erfc_r = f2(b2_erfc(ar.v.x), b2_erfc(ar.v.y))*rinv;
            const F2 ar2 = ar*ar;
            ec = kqq*erfc_r;
            rfc = kqq*fma2(f2(p.tasp), f2(b2_exp(-ar2.v.x), b2_exp(-ar2.v.y)), erfc_r);
        }
        F2 rF;
        if (SW == SW_NONE) {
            rF = rflj + rfc;
        } else {
            // switching function S(u) = 1 - u^3 (10 - 15 u + 6 u^2) and r dS/dr = -30 u^2 (1-u)^2 r du/dr
            const F2 r = r2*rinv;
            F2 u, rdu;
            if (SWF == SWF_LINEAR) {
                u = max0(fma2(r, f2(p.iw), f2(-p.rs*p.iw)));
                rdu = r*f2(p.iw);
            } else {
                F2 rd = r2;
                for (int k = 2; k < p.degree; k++) rd = rd*r;
                u = max0(fma2(rd, f2(p.iwd), f2(-p.rsd*p.iwd)));
                rdu = rd*f2((float)p.degree*p.iwd);
            }
            const F2 u2 = u*u;
            const F2 S = fma2(u2*u, fma2(fma2(u, f2(-6.f), f2(15.f)), u, f2(-10.f)), f2(1.f));
            if (SW == SW_ALL && VAR == VAR_FSWITCH) {
                rF = S*(rflj + rfc);
            } else {
                const F2 om = fma2(u, f2(-1.f), f2(1.f));
                const F2 rdS = (u2*f2(-30.f))*(om*om)*rdu;
                if (SW == SW_LJ) {
                    rF = fma2(S, rflj, rfc) + (rdS*elj)*f2(-1.f);
                } else if (VAR == VAR_NONE) {
                    rF = fma2(S, rflj + rfc, (rdS*(elj + ec))*f2(-1.f));
                } else {      // VAR_SHIFT: V - V(rc0)
                    const float irc = p.inv_rc0;
                    const F2 c2 = (sig*sig)*f2(irc*irc);
                    const F2 c6 = c2*c2*c2;
                    // eps here is EPS_SCALE * eps: 4 eps c6 (c6 - 1) = (eps/6) c6 (c6 - 1) for the standard form
                    const F2 vshift = fma2(kqq, f2(irc), ((eps*f2(4.f/EPS_SCALE))*c6)*(c6 + f2(-1.f)));
                    const F2 V = elj + ec + vshift*f2(-1.f);
                    rF = fma2(S, rflj + rfc, (rdS*V)*f2(-1.f));
                }
            }
        }
        return (rF*rinv2)*f2(p.sign);
    }
};

// scalar fall-back with the same interface (soft core: divisions and per-launch couplings, not worth packing)
struct SoftcoreForce2 {
    SoftcorePot<float> pot;
    static constexpr float EPS_SCALE = 1.f;
    B2_HD float charge_scale() const { return 1.f; }
#ifdef __CUDACC__
    __device__ __forceinline__ F2 operator()(F2 r2, F2 qq, F2 sig, F2 eps) const {
        float rF0, rF1, e, ri0, ri1;
        pot.template operator()<false>(r2.v.x, qq.v.x, sig.v.x, eps.v.x, rF0, e, ri0);
        pot.template operator()<false>(r2.v.y, qq.v.y, sig.v.y, eps.v.y, rF1, e, ri1);
        return f2(rF0*ri0, rF1*ri1);
    }
#endif
};

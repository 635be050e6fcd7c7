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

enum { COUL_NONE = 0, COUL_PLAIN = 1, COUL_RF = 2, COUL_ERFC = 3 };
enum { LJ_STD = 0, LJ_VIRIAL = 1 };
enum { SW_NONE = 0, SW_ALL = 1, SW_LJ = 2 };
enum { SWF_LINEAR = 0, SWF_POWER = 1 };
enum { VAR_NONE = 0, VAR_SHIFT = 1, VAR_FSWITCH = 2 };

__device__ __forceinline__ float b2_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double b2_rsqrt(double x) { return 1.0/sqrt(x); }
// 1/r and 1/r^2 from r^2.  fp32: MUFU.RSQ (<= 2 ulp) refined by one Newton step (~0.5 ulp), because the
// r^-12 term amplifies the relative error of 1/r twelve-fold.
__device__ __forceinline__ void b2_inverse(float r2, float& rinv, float& rinv2) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(r2));   // r2 is never subnormal: no range fix-up code
    const float e = fmaf(-r2*y, y, 1.0f);
    y = fmaf(0.5f*y, e, y);
    rinv = y;
    rinv2 = y*y;
}
__device__ __forceinline__ void b2_inverse(double r2, double& rinv, double& rinv2) {
    rinv = 1.0/sqrt(r2);
    rinv2 = 1.0/r2;
}
__device__ __forceinline__ float b2_erfc(float x) { return erfcf(x); }
__device__ __forceinline__ double b2_erfc(double x) { return erfc(x); }
__device__ __forceinline__ float b2_exp(float x) { return __expf(x); }
__device__ __forceinline__ double b2_exp(double x) { return exp(x); }
__device__ __forceinline__ float b2_log(float x) { return __logf(x); }
__device__ __forceinline__ double b2_log(double x) { return log(x); }
__device__ __forceinline__ float b2_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double b2_max(double a, double b) { return fmax(a, b); }

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
__device__ __forceinline__ void switch_eval(const PotParams<T>& p, int swf, T r, T r2, T& S, T& rdS) {
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
    __device__ __forceinline__ void operator()(T r2, T qq, T sig, T eps, T& rF, T& e, T& rinv2) const {
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
    __device__ __forceinline__ void operator()(T r2, T qq, T sig, T eps, T& rF, T& e, T& rinv2) const {
        T dv, dc;
        eval(r2, qq, sig, eps, rF, e, dv, dc);
        rinv2 = T(1)/r2;
    }

    __device__ __forceinline__ void eval(T r2, T qq, T sig, T eps, T& rF, T& e, T& dEdlv, T& dEdlc) const {
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

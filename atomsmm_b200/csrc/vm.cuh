// Device-side stack VM for integrator expressions and custom bonded energies, plus the
// counter-based RNG (Philox4x32-10) behind `gaussian` / `uniform`.
//
// Replaces OpenMM's Lepton interpreter for CustomIntegrator ComputeGlobal / ComputePerDof /
// ComputeSum steps (reference call sites: integrators.py:113,129,145) and CustomBondForce /
// CustomAngleForce energies (forces.py:338, systems.py:168,228,914,923).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "program.h"

// ---- Philox4x32-10 ----------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0*c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1*c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}

__device__ __forceinline__ void philox4x32(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t (&out)[4]) {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int i = 0; i < 10; i++) philox_round(c, k);
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

__device__ __forceinline__ double u32x2_to_unit(uint32_t a, uint32_t b) {
    // 53-bit uniform in (0,1)
    uint64_t x = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11);
    return ((double)(x & ((1ull << 53) - 1)) + 0.5)*(1.0/9007199254740992.0);
}

struct RngStream {
    uint64_t seed;
    uint32_t c0, c1, c2;   // (element, step, op serial)
    uint32_t draw;
    __device__ double uniform() {
        uint32_t r[4];
        philox4x32(seed, c0, c1, c2, draw++, r);
        return u32x2_to_unit(r[0], r[1]);
    }
    __device__ double gaussian() {
        uint32_t r[4];
        philox4x32(seed, c0, c1, c2, draw++, r);
        double u1 = u32x2_to_unit(r[0], r[1]), u2 = u32x2_to_unit(r[2], r[3]);
        return sqrt(-2.0*log(u1))*cospi(2.0*u2);
    }
};

// ---- variable access ----------------------------------------------------------------------------
struct PerDofTable {
    double* vars[2 + B2_MAX_PERDOF];   // 0 x, 1 v, 2.. user
    const float4* f[33];
    const double* mass;
};

struct VmScalarEnv {       // custom bonded terms: variables come from a small local array
    const double* vars;
};

__device__ __forceinline__ double vm_powi(double x, int n) {
    bool inv = n < 0;
    unsigned m = inv ? -n : n;
    double r = 1.0;
    while (m) {
        if (m & 1) r *= x;
        x *= x;
        m >>= 1;
    }
    return inv ? 1.0/r : r;
}

// Rare / heavy opcodes live in one out-of-line function so that the interpreter's hot loop (pushes,
// + - * and stores) is a few hundred bytes of code: a scalar program is executed by ONE thread, whose
// speed is set by instruction-cache misses, not by arithmetic (measured: ~500 cycles per VM
// instruction when every opcode jumped into a different part of a 30 KB switch).
static __device__ __noinline__ int vm_cold(int op, int arg, double* st, int sp, RngStream* rng) {
    switch (op) {
    case VM_GAUSS: st[sp++] = rng->gaussian(); break;
    case VM_UNIF: st[sp++] = rng->uniform(); break;
    case VM_DIV: sp--; st[sp-1] /= st[sp]; break;
    case VM_POW: sp--; st[sp-1] = pow(st[sp-1], st[sp]); break;
    case VM_POWI: st[sp-1] = vm_powi(st[sp-1], arg); break;
    case VM_SQRT: st[sp-1] = sqrt(st[sp-1]); break;
    case VM_EXP: st[sp-1] = exp(st[sp-1]); break;
    case VM_LOG: st[sp-1] = log(st[sp-1]); break;
    case VM_SIN: st[sp-1] = sin(st[sp-1]); break;
    case VM_COS: st[sp-1] = cos(st[sp-1]); break;
    case VM_TAN: st[sp-1] = tan(st[sp-1]); break;
    case VM_ERF: st[sp-1] = erf(st[sp-1]); break;
    case VM_ERFC: st[sp-1] = erfc(st[sp-1]); break;
    case VM_ABS: st[sp-1] = fabs(st[sp-1]); break;
    case VM_MIN: sp--; st[sp-1] = fmin(st[sp-1], st[sp]); break;
    case VM_MAX: sp--; st[sp-1] = fmax(st[sp-1], st[sp]); break;
    case VM_STEP: st[sp-1] = st[sp-1] < 0.0 ? 0.0 : 1.0; break;
    case VM_DELTA: st[sp-1] = st[sp-1] == 0.0 ? 1.0 : 0.0; break;
    case VM_SELECT: sp -= 2; st[sp-1] = (st[sp-1] != 0.0) ? st[sp] : st[sp+1]; break;
    case VM_FLOOR: st[sp-1] = floor(st[sp-1]); break;
    case VM_CEIL: st[sp-1] = ceil(st[sp-1]); break;
    default: break;
    }
    return sp;
}

// MODE 0: per-DOF (tab/dof valid)   MODE 1: scalar program over globals (STOREG/JMP allowed)
// MODE 2: custom bonded term (PUSHV reads env.vars)
template <int MODE>
__device__ double vm_run(const int* __restrict__ code, int len, const double* __restrict__ consts,
                         double* globals, const PerDofTable* tab, int dof, const double* lvars,
                         RngStream* rng, const double* energies) {
    double st[B2_VM_STACK];
    int sp = 0;
    int pc = 0;
    while (pc < len) {
        const int op = code[2*pc], arg = code[2*pc + 1];
        pc++;
        if (op == VM_PUSHC) st[sp++] = consts[arg];
        else if (op == VM_PUSHG) st[sp++] = globals[arg];
        else if (op == VM_MUL) { sp--; st[sp-1] *= st[sp]; }
        else if (op == VM_ADD) { sp--; st[sp-1] += st[sp]; }
        else if (op == VM_SUB) { sp--; st[sp-1] -= st[sp]; }
        else if (op == VM_NEG) st[sp-1] = -st[sp-1];
        else if (op == VM_PUSHV) st[sp++] = MODE == 0 ? tab->vars[arg][dof] : lvars[arg];
        else if (op == VM_STOREG) { if (MODE == 1) globals[arg] = st[--sp]; }
        else if (op == VM_PUSHM) st[sp++] = tab->mass[dof/3];
        else if (op == VM_PUSHF) {
            const float* f = reinterpret_cast<const float*>(tab->f[arg]);
            st[sp++] = (double)f[(dof/3)*4 + dof%3];
        }
        else if (op == VM_PUSHE) st[sp++] = energies[arg];
        else if (op == VM_JMP) { if (MODE == 1) pc = arg; }
        else if (op == VM_JMPZ) { if (MODE == 1) { if (st[--sp] == 0.0) pc = arg; } }
        else if (op == VM_CMP) {
            sp--;
            const double a = st[sp-1], b = st[sp];
            const bool r = arg == 0 ? a == b : arg == 1 ? a < b : arg == 2 ? a > b : arg == 3 ? a != b
                         : arg == 4 ? a <= b : a >= b;
            st[sp-1] = r ? 1.0 : 0.0;
        }
        else sp = vm_cold(op, arg, st, sp, rng);
    }
    return sp > 0 ? st[sp-1] : 0.0;
}

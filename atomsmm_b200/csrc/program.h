// Lowered integrator program: op records and VM opcodes shared by host and device code.
// The Python front end (atomsmm_b200/lowering.py, expr.py) emits exactly these numbers.
#pragma once

struct b2_op {
    int kind;
    int a, b, c, d, e, f, g;
};

enum {
    // target[dof] = eval(code) for every degree of freedom.
    //   a: target variable (0 x, 1 v, 2+k user per-DOF k)   b: code offset (ints)   c: code length
    //   d: 1 if the expression draws random numbers          e: op serial (RNG stream id)
    B2_OP_PERDOF = 1,
    // globals[a] = sum over DOFs of eval(code).   b, c as above
    B2_OP_SUM = 2,
    // scalar program (assignments to globals, if/while on globals).   b, c as above
    B2_OP_GLOBAL = 3,
    // make force slot b valid for the current positions.   a: force-group mask
    B2_OP_EVAL = 4,
    // v += sum_k sign_k g[coef_k] f[slot_k] / m, optionally followed by x += g[c] v   (fused
    // kick[+drift]).  a: number of terms (<= B2_MAX_KICK_TERMS), b: offset in the code pool of the
    // term table {slot, coef, sign} x a, c: global index of the drift coefficient or -1
    B2_OP_KICK = 5,
    // x += g[a] * v                          fast drift
    B2_OP_DRIFT = 6,
    // v *= g[a]                              fast scale
    B2_OP_SCALE = 7,
    // UpdateContextState hook: where a MonteCarloBarostat of the System acts.  Executed by program_run
    // BETWEEN steps (the hook is the first computation of every atomsmm step program, integrators.py:115-122)
    B2_OP_UPDATE_STATE = 8,
    // dE/d(parameter) of the soft-core pair forces -> device energy slots 64 (lambda_vdw) and 65
    // (lambda_coul), read by VM_PUSHE: the `deriv(energy, lambda)` of AFED (integrators.py:735-737)
    B2_OP_ENERGY = 9,
    // fused RESPA inner loop: see integrate.cu.  a: iterations, b: g-index of kick coefficient,
    //   c: g-index of drift coefficient, d: force slot of the inner group
    B2_OP_FUSED_INNER = 10,
    // a scalar program has moved a context parameter that forces depend on: every cached force is stale
    B2_OP_INVALIDATE = 11,
    // CustomIntegrator.addConstrainPositions / addConstrainVelocities (SHAKE / RATTLE, constraint.cu)
    B2_OP_CONSTRAIN_X = 12,
    B2_OP_CONSTRAIN_V = 13,
    // hint: the NEXT op is a pure rescaling v <- s v after which  globals[a] * globals[b]^2  is the sum
    // m v.v of the new velocities (a thermostat chain keeps every factor but the last in globals[a]): the engine
    // carries the sum to the next thermostat block instead of sweeping v (and, with several ranks, reducing) again
    B2_OP_MVV_FACTOR = 14,
};

// VM opcodes (two ints per instruction: opcode, argument)
enum {
    VM_PUSHC = 0, VM_PUSHG = 1, VM_PUSHV = 2, VM_GAUSS = 3, VM_UNIF = 4, VM_ADD = 5, VM_SUB = 6,
    VM_MUL = 7, VM_DIV = 8, VM_NEG = 9, VM_POW = 10, VM_POWI = 11, VM_SQRT = 12, VM_EXP = 13,
    VM_LOG = 14, VM_SIN = 15, VM_COS = 16, VM_TAN = 17, VM_ERF = 18, VM_ERFC = 19, VM_ABS = 20,
    VM_MIN = 21, VM_MAX = 22, VM_STEP = 23, VM_DELTA = 24, VM_SELECT = 25, VM_FLOOR = 26,
    VM_CEIL = 27, VM_PUSHM = 28, VM_PUSHF = 29, VM_DERIV = 30, VM_STOREG = 31, VM_JMP = 32,
    VM_JMPZ = 33, VM_CMP = 34, VM_PUSHE = 35,
};

#define B2_MAX_KICK_TERMS 6
#define B2_MAX_PERDOF 14
#define B2_VM_STACK 24

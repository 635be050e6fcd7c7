"""
Run-time compilation of generic per-DOF expressions (csrc/jit.cu): ComputePerDof / ComputeSum
steps that are not a recognised kick, drift or rescaling -- massive thermostats, noise terms -- are translated from
the engine's bytecode into CUDA C and compiled by NVRTC for sm_100a.  NVRTC needs no GPU, so the translator and the
compilation are checked here for every per-DOF and sum expression the reference's integrators lower to
(b2_jit_check, a pure host function of the C library); the execution is checked on the GPU in
tests/test_gpu_variants.py (identical bits with the compiler on and off, and against the float64 oracle).
"""

import ctypes

import numpy as np
import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import engine, lowering, unit

import systems
from test_lowered_programs import CASES, water_cluster

OP_PERDOF, OP_SUM = 1, 2


def _check(code):
    lib = engine.library()
    lib.b2_jit_check.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    lib.b2_jit_check.restype = ctypes.c_int
    code = np.ascontiguousarray(code, dtype=np.int32)
    out = ctypes.create_string_buffer(1 << 16)
    status = lib.b2_jit_check(code.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(code)//2, out, len(out))
    return status, out.value.decode()


def test_translation_of_a_langevin_type_expression():
    # v*exp(-g) + sqrt(c/m)*gaussian, written as the engine's stack bytecode
    PUSHC, PUSHG, PUSHV, GAUSS, ADD, MUL, DIV, NEG, SQRT, EXP, PUSHM = 0, 1, 2, 3, 5, 7, 8, 9, 12, 13, 28
    program = [(PUSHV, 1), (PUSHG, 3), (NEG, 0), (EXP, 0), (MUL, 0), (PUSHC, 0), (PUSHM, 0), (DIV, 0), (SQRT, 0),
               (GAUSS, 0), (MUL, 0), (ADD, 0)]
    status, text = _check(np.array(program).ravel())
    if status == -3 and 'libnvrtc' in text:
        pytest.skip('libnvrtc is not installed')
    assert status == 0, text
    lines = [line.strip() for line in text.strip().split('\n')]
    assert lines[0] == 'const double t0 = tab.vars[1][dof];'
    assert 'const double t3 = exp(t2);' in lines
    assert 'const double t9 = rng.gaussian();' in lines          # random draws stay in bytecode order
    assert lines[-2] == 'const double t11 = t4 + t10;'
    assert 'sm_100a' in lines[-1]


def test_scalar_program_opcodes_are_refused():
    STOREG, PUSHC = 31, 0
    status, _ = _check(np.array([(PUSHC, 0), (STOREG, 1)]).ravel())
    assert status == -3 or status == 0 and False


@pytest.mark.parametrize('case', ['nhl-r-massive', 'mts-middle-nres2', 'respa-memory'])
def test_every_generic_step_of_the_reference_integrators_compiles(case):
    system, pos = water_cluster(24)
    respa = atomsmm.RESPASystem(system, 7*unit.angstroms, 5*unit.angstroms)
    dof = atomsmm.countDegreesOfFreedom(respa)
    integrator = CASES[case](dof)
    program = lowering.lower_program(integrator, fast=False)      # fast=False: every per-DOF step is generic
    code = np.array(program.bc.code, dtype=np.int32)
    checked = 0
    for op in program.ops:
        kind, a, b, c, d = op[:5]
        if kind == OP_PERDOF or (kind == OP_SUM and d != 1):
            status, text = _check(code[b:b + 2*c])
            if status == -3 and 'libnvrtc' in text:
                pytest.skip('libnvrtc is not installed')
            assert status == 0, text
            checked += 1
    assert checked > 0


# ---- semantics of the translator, every opcode: generated statements run on the host against a Python stack machine --
_OPS = dict(PUSHC=0, PUSHG=1, PUSHV=2, GAUSS=3, UNIF=4, ADD=5, SUB=6, MUL=7, DIV=8, NEG=9, POW=10, POWI=11, SQRT=12, EXP=13,
            LOG=14, SIN=15, COS=16, TAN=17, ERF=18, ERFC=19, ABS=20, MIN=21, MAX=22, STEP=23, DELTA=24, SELECT=25, FLOOR=26,
            CEIL=27, PUSHM=28, PUSHF=29, CMP=34)
_PUSH = ['PUSHC', 'PUSHG', 'PUSHV', 'GAUSS', 'UNIF', 'PUSHM', 'PUSHF']
_UNARY = ['NEG', 'POWI', 'SQRT', 'EXP', 'LOG', 'SIN', 'COS', 'TAN', 'ERF', 'ERFC', 'ABS', 'STEP', 'DELTA', 'FLOOR', 'CEIL']
_BINARY = ['ADD', 'SUB', 'MUL', 'DIV', 'POW', 'MIN', 'MAX', 'CMP']


def _random_program(rng, length):
    """A well-formed postfix program (stack never underflows, one value left)."""
    program, depth = [], 0
    while len(program) < length or depth != 1:
        choices = []
        if depth < 6 and len(program) < length:
            choices.append('push')
        if depth >= 1 and len(program) < length:
            choices.append('unary')
        if depth >= 2:
            choices.append('binary')
        if depth >= 3:
            choices.append('select')
        kind = choices[rng.integers(len(choices))]
        if kind == 'push':
            name = _PUSH[rng.integers(len(_PUSH))]
            arg = int(rng.integers(4)) if name in ('PUSHC', 'PUSHG') else int(rng.integers(3)) if name in ('PUSHV', 'PUSHF') else 0
            depth += 1
        elif kind == 'unary':
            name = _UNARY[rng.integers(len(_UNARY))]
            arg = int(rng.integers(-3, 5)) if name == 'POWI' else 0
        elif kind == 'binary':
            name = _BINARY[rng.integers(len(_BINARY))]
            arg = int(rng.integers(6)) if name == 'CMP' else 0
            depth -= 1
        else:
            name, arg = 'SELECT', 0
            depth -= 2
        program.append((name, arg))
    return program


def _python_vm(program, env, draws):
    """The interpreter of csrc/vm.cuh restated in Python (float64, same operand order)."""
    import math
    stack, draws = [], list(draws)

    def safe(fn, *args):
        try:
            return float(fn(*args))
        except (ValueError, OverflowError, ZeroDivisionError):
            return float('nan')
    for name, arg in program:
        if name == 'PUSHC': stack.append(env['consts'][arg])
        elif name == 'PUSHG': stack.append(env['globals'][arg])
        elif name == 'PUSHV': stack.append(env['vars'][arg])
        elif name == 'PUSHM': stack.append(env['mass'])
        elif name == 'PUSHF': stack.append(float(np.float32(env['f'][arg])))
        elif name in ('GAUSS', 'UNIF'): stack.append(draws.pop(0))
        elif name == 'NEG': stack.append(-stack.pop())
        elif name == 'POWI':
            x, n = stack.pop(), arg
            r, m = 1.0, abs(n)
            with np.errstate(all='ignore'):
                x = np.float64(x); r = np.float64(1.0)
                while m:
                    if m & 1: r = r*x
                    x = x*x
                    m >>= 1
                stack.append(float(np.float64(1.0)/r if n < 0 else r))
        elif name in ('SQRT', 'EXP', 'LOG', 'SIN', 'COS', 'TAN', 'ABS', 'FLOOR', 'CEIL'):
            fn = dict(SQRT=np.sqrt, EXP=np.exp, LOG=np.log, SIN=np.sin, COS=np.cos, TAN=np.tan, ABS=np.abs, FLOOR=np.floor,
                      CEIL=np.ceil)[name]
            with np.errstate(all='ignore'):
                stack.append(float(fn(np.float64(stack.pop()))))
        elif name == 'ERF': stack.append(safe(math.erf, stack.pop()))
        elif name == 'ERFC': stack.append(safe(math.erfc, stack.pop()))
        elif name == 'STEP': stack.append(0.0 if stack.pop() < 0.0 else 1.0)
        elif name == 'DELTA': stack.append(1.0 if stack.pop() == 0.0 else 0.0)
        elif name == 'SELECT':
            c, b, a = stack.pop(), stack.pop(), stack.pop()
            stack.append(b if a != 0.0 else c)
        else:
            b, a = stack.pop(), stack.pop()
            with np.errstate(all='ignore'):
                a64, b64 = np.float64(a), np.float64(b)
                if name == 'ADD': r = a64 + b64
                elif name == 'SUB': r = a64 - b64
                elif name == 'MUL': r = a64*b64
                elif name == 'DIV': r = a64/b64
                elif name == 'POW': r = np.power(a64, b64)
                elif name == 'MIN': r = np.fmin(a64, b64)
                elif name == 'MAX': r = np.fmax(a64, b64)
                else: r = [a == b, a < b, a > b, a != b, a <= b, a >= b][arg] and 1.0 or 0.0
            stack.append(float(r))
    return stack[-1]


def test_generated_statements_have_the_semantics_of_the_interpreter(tmp_path):
    """Random well-formed programs over EVERY per-DOF opcode: the CUDA C statements csrc/jit.cu generates are compiled
    for the host (g++) around stand-ins for the device tables and a scripted random stream, and must reproduce a Python
    restatement of the bytecode interpreter -- operand order of - / pow CMP SELECT, integer powers, order of the random
    draws."""
    import shutil
    import subprocess
    if shutil.which('g++') is None:
        pytest.skip('g++ not available')
    rng = np.random.default_rng(2024)
    programs = [_random_program(rng, int(rng.integers(3, 28))) for _ in range(120)]
    env = dict(consts=[0.75, -1.5, 2.0, 0.3], globals=[1.25, -0.5, 3.0, 0.125], vars=[0.4, -1.1, 2.2], mass=15.999,
               f=[12.5, -7.25, 0.5])
    draws = [0.3, -1.2, 0.8, 0.05, -0.6, 1.7, 0.45, -0.1, 2.1, -0.9]*6
    functions = []
    for k, program in enumerate(programs):
        code = np.array([(_OPS[name], arg) for name, arg in program], dtype=np.int32).ravel()
        status, text = _check(code)
        if status == -3 and 'libnvrtc' in text:
            pytest.skip('libnvrtc is not installed')
        assert status == 0, (program, text)
        lines = text.rstrip().split('\n')
        result = lines[-1].split('->')[1].split()[0]
        functions.append('static double eval_%d(int dof, Tab tab, const double* consts, const double* globals) {\n'
                         '    Rng rng;\n%s\n    return %s;\n}\n' % (k, '\n'.join(lines[:-1]), result))
    source = '''
#include <cmath>
#include <cstdio>
struct float4 { float x, y, z, w; };
struct Tab { double* vars[16]; const float4* f[33]; const double* mass; };
static const double DRAWS[] = {%s};
struct Rng { int n = 0; double gaussian() { return DRAWS[n++]; } double uniform() { return DRAWS[n++]; } };
static double vm_powi(double x, int n) { bool inv = n < 0; unsigned m = inv ? -n : n; double r = 1.0;
    while (m) { if (m & 1) r *= x; x *= x; m >>= 1; } return inv ? 1.0/r : r; }
%s
int main() {
    double v0[3] = {%r, 0, 0}, v1[3] = {%r, 0, 0}, v2[3] = {%r, 0, 0}, mass[1] = {%r};
    float4 f0[1] = {{%rf, 0, 0, 0}}, f1[1] = {{%rf, 0, 0, 0}}, f2[1] = {{%rf, 0, 0, 0}};
    double consts[4] = {%s}, globals[4] = {%s};
    Tab tab = {}; tab.vars[0] = v0; tab.vars[1] = v1; tab.vars[2] = v2; tab.f[0] = f0; tab.f[1] = f1; tab.f[2] = f2; tab.mass = mass;
%s
    return 0;
}
''' % (', '.join(repr(d) for d in draws), '\n'.join(functions), env['vars'][0], env['vars'][1], env['vars'][2], env['mass'],
       env['f'][0], env['f'][1], env['f'][2], ', '.join(repr(c) for c in env['consts']), ', '.join(repr(g) for g in env['globals']),
       '\n'.join('    printf("%%.17g\\n", eval_%d(0, tab, consts, globals));' % k for k in range(len(programs))))
    path = tmp_path/'generated.cpp'
    path.write_text(source)
    exe = str(tmp_path/'generated')
    subprocess.run(['g++', '-O0', '-ffp-contract=off', '-o', exe, str(path)], check=True, capture_output=True)
    got = [float(x) for x in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    assert len(got) == len(programs)
    for program, value in zip(programs, got):
        want = _python_vm(program, env, draws)
        if np.isnan(want):
            assert np.isnan(value), program
        elif np.isinf(want):
            assert value == want, program
        else:
            assert value == pytest.approx(want, rel=1e-12, abs=1e-300), program

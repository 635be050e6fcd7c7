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

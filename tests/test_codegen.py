"""
Groundwork for compiled scalar programs (atomsmm_b200/codegen.py): the C text generated
from the VM bytecode of a lowered integrator program is compiled here as host C with gcc and must
compute exactly what the bytecode computes (reference execution: tests/lowered_executor.py's VM), for
the chained Nose-Hoover program of BASELINE config 2, Bussi's rejection loop (while / if blocks, random
numbers), the AFED wall reflection (select / step) and a Nose-Hoover sub-loop.  The engine does not use
this path yet.
"""

import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import codegen, lowering, unit

from lowered_executor import Executor

fs, ps, K = unit.femtoseconds, unit.picoseconds, unit.kelvin

RNG_STUB = '''
struct b2_rng { const double* values; int next; };
double b2_rng_gaussian(b2_rng* rng) { return rng->values[rng->next++]; }
double b2_rng_uniform(b2_rng* rng) { return rng->values[rng->next++]; }
'''


class Tape(object):
    """Random numbers from a fixed tape, for the C stub and the Python VM alike."""

    def __init__(self, values):
        self.values, self.next = values, 0

    def standard_normal(self, *args):
        self.next += 1
        return self.values[self.next - 1]

    random = standard_normal


def compile_programs(program):
    sources = codegen.program_sources(program)
    text = codegen.PRELUDE + RNG_STUB + '\n'.join(sources.values())
    wrappers = []
    for index in sources:
        wrappers.append('void run_%d(double* G, const double* E, const double* tape) {\n'
                        '    struct b2_rng rng = {tape, 0};\n    b2_scalar_program_%d(G, E, &rng);\n}' % (index, index))
    folder = tempfile.mkdtemp(prefix='b2codegen')
    path = os.path.join(folder, 'programs.c')
    with open(path, 'w') as handle:
        handle.write(text + '\n' + '\n'.join(wrappers) + '\n')
    library = os.path.join(folder, 'programs.so')
    subprocess.run(['gcc', '-O2', '-shared', '-fPIC', '-o', library, path, '-lm'], check=True)
    return ctypes.CDLL(library), sources


def check(integrator, parameters=None, derivative_slots=None, trials=5, constrained=False):
    program = lowering.lower_program(integrator, 0xffffffff, parameters or {}, True, constrained=constrained,
                                     derivative_slots=derivative_slots)
    lib, sources = compile_programs(program)
    assert sources
    rng = np.random.default_rng(3)
    vm = Executor.__new__(Executor)                 # only the VM of the executor is needed
    vm.code, vm.consts = list(program.bc.code), np.array(program.bc.consts if program.bc.consts else [0.0])
    vm.param_names, vm.parameters, vm.n = {}, {}, 1
    for index in sources:
        op = program.ops[index]
        start, length = (op[2], op[3]) if op[0] == lowering.OP_GLOBAL else (op[6], op[7])
        for _ in range(trials):
            values = np.array(program.global_values, dtype=np.float64)
            values *= 1.0 + 0.3*rng.standard_normal(len(values))
            values[values == 0] = rng.uniform(0.1, 2.0, size=int(np.sum(values == 0)))
            if 'mvv' in program.global_names and 'LkT' in program.global_names:
                values[program.gindex('mvv')] = abs(values[program.gindex('LkT')])*rng.uniform(0.8, 1.2)
            tape = rng.uniform(0.05, 0.95, size=4096)
            energies = rng.standard_normal(96)
            vm.globals, vm.energies, vm.rng = values.copy(), energies, Tape(tape)
            vm.run_vm(start, length)
            compiled = values.copy()
            getattr(lib, 'run_%d' % index)(compiled.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                           energies.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                           tape.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
            assert np.allclose(compiled, vm.globals, rtol=1e-13, atol=1e-300), (index, compiled - vm.globals)
    return program, sources


def test_chained_nose_hoover_program_compiles_to_straight_line_c():
    nh = atomsmm.NoseHooverPropagator(300*K, 4605, 100*fs)
    integrator = atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([4, 2, 1]),
                                                 atomsmm.SuzukiYoshidaPropagator(nh, 3)).integrator(4*fs)
    program, sources = check(integrator)
    assert all('goto' not in text for text in sources.values())         # no control flow: straight-line code
    assert max(text.count('\n') for text in sources.values()) > 100    # the 124-instruction chain


def test_bussi_rejection_loop_and_nose_hoover_subloop():
    thermostat = atomsmm.VelocityRescalingPropagator(300*K, 4605, 0.1*ps)
    integrator = atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([2, 1]), thermostat).integrator(1*fs)
    program, sources = check(integrator, trials=12)
    assert any('goto' in text for text in sources.values())             # while / if blocks became gotos
    looped = atomsmm.TrotterSuzukiPropagator(atomsmm.propagators.UnconstrainedVelocityVerletPropagator(),
                                             atomsmm.NoseHooverPropagator(300*K, 4605, 100*fs, 4)).integrator(1*fs)
    check(looped)


def test_afed_wall_reflection_and_derivative():
    nvt = atomsmm.TrotterSuzukiPropagator(atomsmm.VelocityVerletPropagator(),
                                          atomsmm.NoseHooverPropagator(300*K, 4491, 10*fs)).integrator(1*fs)
    variable = atomsmm.ExtendedSystemVariable('lambda_vdw', 1000, 5, 40*fs)
    integrator = atomsmm.AdiabaticDynamicsIntegrator(nvt, 2, [variable])
    program, sources = check(integrator, parameters={'lambda_vdw': 1.0},
                             derivative_slots={'lambda_vdw': lowering.ENERGY_SLOT_DLAMBDA_VDW}, trials=12)
    assert any('E[64]' in text for text in sources.values())           # deriv(energy, lambda_vdw)


def test_unsupported_opcode_is_reported():
    with pytest.raises(codegen.CodegenError):
        codegen.scalar_program_source([2, 0, 31, 0], 0, 2, [0.0])      # PUSHV is a per-DOF opcode


def test_generated_source_is_valid_cuda_device_code():
    """The same text, with B2_DEVICE = __device__, cross-compiles for sm_100a (what an NVRTC path will do)."""
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    nh = atomsmm.NoseHooverPropagator(300*K, 4605, 100*fs)
    integrator = atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([4, 2, 1]),
                                                 atomsmm.SuzukiYoshidaPropagator(nh, 3)).integrator(4*fs)
    program = lowering.lower_program(integrator, 0xffffffff, {}, True, constrained=False)
    sources = codegen.program_sources(program)
    device_rng = ('struct b2_rng { unsigned long long state; };\n'
                  '__device__ double b2_rng_gaussian(b2_rng* rng) { return (double)(rng->state++ & 7); }\n'
                  '__device__ double b2_rng_uniform(b2_rng* rng) { return (double)(rng->state++ & 7)*0.125; }\n')
    kernels = ''.join('__global__ void k_%d(double* G, const double* E) { b2_rng rng = {0}; '
                      'if (threadIdx.x == 0) b2_scalar_program_%d(G, E, &rng); }\n' % (k, k) for k in sources)
    text = '#define B2_DEVICE __device__\n' + codegen.PRELUDE + device_rng + '\n'.join(sources.values()) + '\n' + kernels
    folder = tempfile.mkdtemp(prefix='b2codegen')
    path = os.path.join(folder, 'programs.cu')
    with open(path, 'w') as handle:
        handle.write(text)
    result = subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-c', path, '-o',
                             os.path.join(folder, 'programs.o')], capture_output=True, text=True)
    assert result.returncode == 0, result.stderr

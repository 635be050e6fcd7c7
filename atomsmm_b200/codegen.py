"""
Source generation for scalar integrator programs (groundwork for SURVEY 8f rank 4; NOT used by the
engine yet -- the CUDA engine still interprets these programs with its stack VM, csrc/vm.cuh).

The interpreter costs ~130 ns per VM instruction on one GPU thread (a chain of
three Nose-Hoover blocks is 124 instructions, 16 us per reduction kernel, ~4 % of a BASELINE config-2
step).  Scalar programs are tiny and fixed once an integrator is built, so they can be turned into
straight-line C: ``scalar_program_source`` translates the VM bytecode of one program (the same
bytecode the engine receives, csrc/program.h) into a C function over the global-variable array.
Stack depths are static in this bytecode (it is compiled from expression trees; jumps happen at depth
0 or right after a comparison), so the VM stack becomes local variables and if/while blocks become
gotos.  The text compiles unchanged as host C (tests/test_codegen.py checks it against the reference
execution of the bytecode) and as CUDA device code (``__device__`` is supplied through a macro), which
is what an NVRTC path will feed to the driver.
"""

from . import expr as X

_NAMES = {v: k for k, v in X.OPCODES.items()}
_UNARY = dict(SQRT='sqrt', EXP='exp', LOG='log', SIN='sin', COS='cos', TAN='tan', ERF='erf', ERFC='erfc',
              ABS='fabs', FLOOR='floor', CEIL='ceil')
_BINARY = dict(ADD='+', SUB='-', MUL='*', DIV='/')
_COMPARE = ['==', '<', '>', '!=', '<=', '>=']


class CodegenError(Exception):
    pass


def scalar_program_source(code, start, length, consts, name='b2_scalar_program'):
    """C source of ``void name(double* G, const double* E, b2_rng* rng)`` equivalent to running VM
    instructions [start, start+length) of the int stream ``code`` (pairs opcode, argument) as a scalar
    program: G = integrator globals, E = device energy slots (VM_PUSHE), rng = random stream with
    ``b2_rng_gaussian`` / ``b2_rng_uniform``.  Constants are baked in as literals."""
    instructions = [(_NAMES[code[start + 2*k]], code[start + 2*k + 1]) for k in range(length)]
    targets = {arg for op, arg in instructions if op in ('JMP', 'JMPZ')}
    # static stack depth before every instruction
    depth = [None]*(length + 1)
    depth[0] = 0
    work = [0]
    while work:
        pc = work.pop()
        d = depth[pc]
        if pc == length:
            continue
        op, arg = instructions[pc]
        pops, pushes = _effect(op)
        if d < pops:
            raise CodegenError('stack underflow at instruction %d (%s)' % (pc, op))
        after = d - pops + pushes
        successors = []
        if op == 'JMP':
            successors = [arg]
        elif op == 'JMPZ':
            successors = [arg, pc + 1]
        else:
            successors = [pc + 1]
        for nxt in successors:
            if nxt > length:
                raise CodegenError('jump outside the program')
            if depth[nxt] is None:
                depth[nxt] = after
                work.append(nxt)
            elif depth[nxt] != after:
                raise CodegenError('inconsistent stack depth at instruction %d' % nxt)
    maxdepth = max(d for d in depth if d is not None) + 1
    lines = ['B2_DEVICE void %s(double* G, const double* E, b2_rng* rng) {' % name,
             '    double %s;' % ', '.join('s%d = 0.0' % k for k in range(maxdepth)),
             '    (void)E; (void)rng; %s' % ' '.join('(void)s%d;' % k for k in range(maxdepth))]
    for pc, (op, arg) in enumerate(instructions):
        if depth[pc] is None:
            continue        # unreachable
        d = depth[pc]
        label = 'L%d: ' % pc if pc in targets else ''
        top, below = 's%d' % (d - 1), 's%d' % (d - 2)
        new = 's%d' % d
        if op == 'PUSHC':
            stmt = '%s = %s;' % (new, _literal(consts[arg]))
        elif op == 'PUSHG':
            stmt = '%s = G[%d];' % (new, arg)
        elif op == 'PUSHE':
            stmt = '%s = E[%d];' % (new, arg)
        elif op == 'GAUSS':
            stmt = '%s = b2_rng_gaussian(rng);' % new
        elif op == 'UNIF':
            stmt = '%s = b2_rng_uniform(rng);' % new
        elif op in _BINARY:
            stmt = '%s = %s %s %s;' % (below, below, _BINARY[op], top)
        elif op == 'NEG':
            stmt = '%s = -%s;' % (top, top)
        elif op == 'POW':
            stmt = '%s = pow(%s, %s);' % (below, below, top)
        elif op == 'POWI':
            stmt = '%s = %s;' % (top, _powi(top, arg))
        elif op in _UNARY:
            stmt = '%s = %s(%s);' % (top, _UNARY[op], top)
        elif op == 'MIN':
            stmt = '%s = fmin(%s, %s);' % (below, below, top)
        elif op == 'MAX':
            stmt = '%s = fmax(%s, %s);' % (below, below, top)
        elif op == 'STEP':
            stmt = '%s = %s < 0.0 ? 0.0 : 1.0;' % (top, top)
        elif op == 'DELTA':
            stmt = '%s = %s == 0.0 ? 1.0 : 0.0;' % (top, top)
        elif op == 'SELECT':
            third = 's%d' % (d - 3)
            stmt = '%s = (%s != 0.0) ? %s : %s;' % (third, third, below, top)
        elif op == 'CMP':
            stmt = '%s = (%s %s %s) ? 1.0 : 0.0;' % (below, below, _COMPARE[arg], top)
        elif op == 'STOREG':
            stmt = 'G[%d] = %s;' % (arg, top)
        elif op == 'JMP':
            stmt = 'goto L%d;' % arg if arg < length else 'return;'
        elif op == 'JMPZ':
            stmt = 'if (%s == 0.0) %s' % (top, 'goto L%d;' % arg if arg < length else 'return;')
        else:
            raise CodegenError('opcode %s has no scalar translation' % op)
        lines.append('    %s%s' % (label, stmt))
    if length in targets:
        lines.append('    L%d: ;' % length)
    lines.append('}')
    return '\n'.join(lines)


def _effect(op):
    """(values popped, values pushed)"""
    if op in ('PUSHC', 'PUSHG', 'PUSHE', 'GAUSS', 'UNIF'):
        return 0, 1
    if op in _BINARY or op in ('POW', 'MIN', 'MAX', 'CMP'):
        return 2, 1
    if op == 'SELECT':
        return 3, 1
    if op in ('STOREG', 'JMPZ'):
        return 1, 0
    if op == 'JMP':
        return 0, 0
    if op in _UNARY or op in ('NEG', 'POWI', 'STEP', 'DELTA'):
        return 1, 1
    raise CodegenError('opcode %s is not allowed in a scalar program' % op)


def _literal(value):
    text = repr(float(value))
    if text in ('inf', '-inf', 'nan'):
        raise CodegenError('non-finite constant')
    return text if ('.' in text or 'e' in text or 'E' in text) else text + '.0'


def _powi(x, n):
    """x^n by repeated squaring, as the VM does (vm_powi)."""
    if n == 0:
        return '1.0'
    m = abs(n)
    squares = [x]
    while (1 << len(squares)) <= m:
        squares.append('(%s*%s)' % (squares[-1], squares[-1]))
    product = '*'.join(sq for bit, sq in enumerate(squares) if m & (1 << bit))
    return '(1.0/(%s))' % product if n < 0 else '(%s)' % product


PRELUDE = '''
#include <math.h>
#ifndef B2_DEVICE
#define B2_DEVICE
#endif
typedef struct b2_rng b2_rng;
B2_DEVICE double b2_rng_gaussian(b2_rng* rng);
B2_DEVICE double b2_rng_uniform(b2_rng* rng);
'''


def program_sources(program):
    """{op index: C source} for every scalar program of a lowered integrator program (stand-alone
    GLOBAL ops and the programs attached to velocity kernels)."""
    from . import lowering as L
    out = {}
    for index, op in enumerate(program.ops):
        if op[0] == L.OP_GLOBAL:
            start, length = op[2], op[3]
        elif op[0] == L.OP_KICK and op[7] > 0:
            start, length = op[6], op[7]
        else:
            continue
        out[index] = scalar_program_source(program.bc.code, start, length, program.bc.consts,
                                           name='b2_scalar_program_%d' % index)
    return out

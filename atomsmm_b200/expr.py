"""
Algebraic-expression front end (host side).

atomsmm describes pair potentials and integrator steps as Lepton-style strings
(reference: forces.py:448-466,539-567; propagators.py:249,271,732-737,1201-1227).  This
module parses those strings into a small AST, inlines the ``;``-separated auxiliary
definitions, evaluates them numerically on the host (used to *recognise* which hand-written
kernel family an energy string belongs to) and compiles them to the stack bytecode executed
by the device-side scalar / per-DOF virtual machine (csrc/vm.cuh).

Grammar: + - * / ^ (right assoc, binds tighter than unary minus), function calls, numbers,
identifiers.  Functions: sqrt exp log sin cos tan erf erfc abs min max step select delta
floor ceil.
"""

import math
import re

_TOKEN = re.compile(r'\s*(?:(\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)|([A-Za-z_][A-Za-z_0-9]*)|(.))')

FUNCTIONS = {
    'sqrt': 1, 'exp': 1, 'log': 1, 'sin': 1, 'cos': 1, 'tan': 1, 'erf': 1, 'erfc': 1,
    'abs': 1, 'step': 1, 'delta': 1, 'floor': 1, 'ceil': 1, 'min': 2, 'max': 2, 'select': 3, 'deriv': 2,
}


class ParseError(ValueError):
    pass


def _tokenize(text):
    tokens = []
    pos = 0
    text = text.strip()
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if m is None:
            raise ParseError('cannot tokenize %r at %d' % (text, pos))
        number, ident, op = m.groups()
        if number is not None:
            tokens.append(('num', float(number)))
        elif ident is not None:
            tokens.append(('id', ident))
        else:
            tokens.append(('op', op))
        pos = m.end()
    tokens.append(('end', None))
    return tokens


class _Parser(object):
    def __init__(self, text):
        self.text = text
        self.tokens = _tokenize(text)
        self.i = 0

    def peek(self):
        return self.tokens[self.i]

    def next(self):
        tok = self.tokens[self.i]
        self.i += 1
        return tok

    def expect(self, op):
        tok = self.next()
        if tok != ('op', op):
            raise ParseError('expected %r in %r' % (op, self.text))

    def parse(self):
        node = self.sum()
        if self.peek()[0] != 'end':
            raise ParseError('trailing tokens in %r' % self.text)
        return node

    def sum(self):
        node = self.product()
        while self.peek() in (('op', '+'), ('op', '-')):
            op = self.next()[1]
            rhs = self.product()
            node = ('add' if op == '+' else 'sub', node, rhs)
        return node

    def product(self):
        node = self.unary()
        while self.peek() in (('op', '*'), ('op', '/')):
            op = self.next()[1]
            rhs = self.unary()
            node = ('mul' if op == '*' else 'div', node, rhs)
        return node

    def unary(self):
        if self.peek() == ('op', '-'):
            self.next()
            return ('neg', self.unary())
        if self.peek() == ('op', '+'):
            self.next()
            return self.unary()
        return self.power()

    def power(self):
        base = self.atom()
        if self.peek() == ('op', '^'):
            self.next()
            exponent = self.unary()  # right associative; allows x^-2
            return ('pow', base, exponent)
        return base

    def atom(self):
        kind, value = self.next()
        if kind == 'num':
            return ('num', value)
        if kind == 'id':
            if self.peek() == ('op', '('):
                self.next()
                args = []
                if self.peek() != ('op', ')'):
                    args.append(self.sum())
                    while self.peek() == ('op', ','):
                        self.next()
                        args.append(self.sum())
                self.expect(')')
                if value not in FUNCTIONS:
                    raise ParseError('unknown function %r in %r' % (value, self.text))
                if FUNCTIONS[value] != len(args):
                    raise ParseError('wrong number of arguments for %s in %r' % (value, self.text))
                return ('call', value, tuple(args))
            return ('var', value)
        if (kind, value) == ('op', '('):
            node = self.sum()
            self.expect(')')
            return node
        raise ParseError('unexpected token %r in %r' % (value, self.text))


def parse_single(text):
    return _Parser(text).parse()


def parse(text):
    """Parse ``main; name1 = expr1; name2 = expr2 ...`` -> (main_ast, {name: ast}).

    A leading ``name=`` on the main term (reference quirk forces.py:472: ``energy=S*(...)``)
    is dropped.  Later definitions may reference earlier or later ones (Lepton semantics).
    """
    parts = [p.strip() for p in text.split(';') if p.strip()]
    if not parts:
        raise ParseError('empty expression')
    main = parts[0]
    m = re.match(r'^([A-Za-z_][A-Za-z_0-9]*)\s*=(?!=)(.*)$', main)
    if m:
        main = m.group(2)
    defs = {}
    for part in parts[1:]:
        m = re.match(r'^([A-Za-z_][A-Za-z_0-9]*)\s*=(?!=)(.*)$', part)
        if not m:
            raise ParseError('bad definition %r' % part)
        name = m.group(1)
        if name not in defs:   # first definition wins (duplicates appear after importFrom twice)
            defs[name] = parse_single(m.group(2))
    return parse_single(main), defs


def substitute(node, defs, _stack=()):
    """Inline auxiliary definitions recursively."""
    kind = node[0]
    if kind == 'num':
        return node
    if kind == 'var':
        name = node[1]
        if name in defs:
            if name in _stack:
                raise ParseError('circular definition of %r' % name)
            return substitute(defs[name], defs, _stack + (name,))
        return node
    if kind == 'call':
        return ('call', node[1], tuple(substitute(a, defs, _stack) for a in node[2]))
    return (kind,) + tuple(substitute(a, defs, _stack) for a in node[1:])


def parse_inlined(text):
    main, defs = parse(text)
    return substitute(main, defs)


def free_symbols(node, out=None):
    if out is None:
        out = set()
    kind = node[0]
    if kind == 'var':
        out.add(node[1])
    elif kind == 'call':
        for a in node[2]:
            free_symbols(a, out)
    elif kind != 'num':
        for a in node[1:]:
            free_symbols(a, out)
    return out


def required_variables(variable, expression):
    """Names an integrator step needs from outside (mirrors integrators.py:91-104)."""
    main, defs = parse(expression)
    symbols = free_symbols(main)
    for d in defs.values():
        free_symbols(d, symbols)
    return sorted(symbols - set(defs) - {variable})


_FUNCS = {
    'sqrt': math.sqrt, 'exp': math.exp, 'log': math.log, 'sin': math.sin, 'cos': math.cos,
    'tan': math.tan, 'erf': math.erf, 'erfc': math.erfc, 'abs': abs, 'min': min, 'max': max,
    'step': lambda x: 0.0 if x < 0 else 1.0,
    'delta': lambda x: 1.0 if x == 0 else 0.0,
    'select': lambda c, a, b: a if c != 0 else b,
    'floor': math.floor, 'ceil': math.ceil,
}


def evaluate(node, env):
    """Evaluate an (inlined) AST with python floats."""
    kind = node[0]
    if kind == 'num':
        return node[1]
    if kind == 'var':
        return env[node[1]]
    if kind == 'add':
        return evaluate(node[1], env) + evaluate(node[2], env)
    if kind == 'sub':
        return evaluate(node[1], env) - evaluate(node[2], env)
    if kind == 'mul':
        return evaluate(node[1], env)*evaluate(node[2], env)
    if kind == 'div':
        return evaluate(node[1], env)/evaluate(node[2], env)
    if kind == 'neg':
        return -evaluate(node[1], env)
    if kind == 'pow':
        return evaluate(node[1], env)**evaluate(node[2], env)
    if kind == 'call':
        if node[1] == 'deriv':   # deriv(energy, parameter): supplied by the caller
            return env['__deriv__'](to_string(node[2][0]), to_string(node[2][1]))
        return _FUNCS[node[1]](*[evaluate(a, env) for a in node[2]])
    raise ParseError('bad node %r' % (node,))


def to_string(node):
    kind = node[0]
    if kind == 'num':
        return repr(node[1])
    if kind == 'var':
        return node[1]
    if kind == 'neg':
        return '(-%s)' % to_string(node[1])
    if kind == 'call':
        return '%s(%s)' % (node[1], ','.join(to_string(a) for a in node[2]))
    sym = {'add': '+', 'sub': '-', 'mul': '*', 'div': '/', 'pow': '^'}[kind]
    return '(%s%s%s)' % (to_string(node[1]), sym, to_string(node[2]))


_COND = re.compile(r'^(.*?)(<=|>=|!=|<|>|=)(.*)$')
COND_OPS = {'=': 0, '<': 1, '>': 2, '!=': 3, '<=': 4, '>=': 5}


def parse_condition(text):
    """``lhs op rhs`` for if/while blocks -> (lhs_ast, opcode, rhs_ast)."""
    depth = 0
    for i, ch in enumerate(text):
        if ch == '(':
            depth += 1
        elif ch == ')':
            depth -= 1
        elif depth == 0 and ch in '<>=!':
            op = text[i:i+2] if text[i:i+2] in ('<=', '>=', '!=') else ch
            if op == '!':
                continue
            lhs, rhs = text[:i], text[i+len(op):]
            return parse_inlined(lhs), COND_OPS[op], parse_inlined(rhs)
    raise ParseError('no comparison operator in condition %r' % text)


# ---------------------------------------------------------------------------------------------
# Bytecode for the device VM (csrc/vm.cuh).  One instruction = (opcode, int arg); constants are
# interned in a double pool.
# ---------------------------------------------------------------------------------------------

OPCODES = dict(
    PUSHC=0, PUSHG=1, PUSHV=2, GAUSS=3, UNIF=4, ADD=5, SUB=6, MUL=7, DIV=8, NEG=9, POW=10, POWI=11,
    SQRT=12, EXP=13, LOG=14, SIN=15, COS=16, TAN=17, ERF=18, ERFC=19, ABS=20, MIN=21, MAX=22,
    STEP=23, DELTA=24, SELECT=25, FLOOR=26, CEIL=27, PUSHM=28, PUSHF=29, DERIV=30, STOREG=31, JMP=32,
    JMPZ=33, CMP=34, PUSHE=35,
)
_CALL_OPS = dict(sqrt='SQRT', exp='EXP', log='LOG', sin='SIN', cos='COS', tan='TAN', erf='ERF',
                 erfc='ERFC', abs='ABS', min='MIN', max='MAX', step='STEP', delta='DELTA',
                 select='SELECT', floor='FLOOR', ceil='CEIL')


class Bytecode(object):
    def __init__(self):
        self.code = []       # flat ints: op, arg, op, arg ...
        self.consts = []

    def emit(self, op, arg=0):
        self.code += [OPCODES[op], int(arg)]

    def const(self, value):
        value = float(value)
        try:
            return self.consts.index(value)
        except ValueError:
            self.consts.append(value)
            return len(self.consts) - 1


VM_STACK = 24      # B2_VM_STACK of csrc/program.h: evaluation stack of the device VM


def stack_depth(node):
    """Largest number of values the RPN code of ``node`` keeps on the VM stack."""
    kind = node[0]
    if kind in ('num', 'var'):
        return 1
    if kind in ('add', 'sub', 'mul', 'div'):
        return max(stack_depth(node[1]), 1 + stack_depth(node[2]))
    if kind == 'neg':
        return stack_depth(node[1])
    if kind == 'pow':
        return max(stack_depth(node[1]), 1 + stack_depth(node[2]))
    if kind == 'call':
        return max([1] + [k + stack_depth(a) for k, a in enumerate(node[2])])
    raise ParseError('bad node %r' % (node,))


def compile_ast(node, resolve, bc, _root=True):
    """Append RPN code for ``node`` to ``bc``.  ``resolve(name)`` -> (op, arg) for a variable.
    An expression deeper than the device VM's stack is refused here (never a silent overflow)."""
    if _root and stack_depth(node) > VM_STACK:
        raise ParseError('expression needs an evaluation stack of %d values; the device VM has %d'
                         % (stack_depth(node), VM_STACK))
    kind = node[0]
    if kind == 'num':
        bc.emit('PUSHC', bc.const(node[1]))
    elif kind == 'var':
        op, arg = resolve(node[1])
        bc.emit(op, arg)
    elif kind in ('add', 'sub', 'mul', 'div'):
        compile_ast(node[1], resolve, bc, False)
        compile_ast(node[2], resolve, bc, False)
        bc.emit(kind.upper())
    elif kind == 'neg':
        compile_ast(node[1], resolve, bc, False)
        bc.emit('NEG')
    elif kind == 'pow':
        compile_ast(node[1], resolve, bc, False)
        exponent = node[2]
        if exponent[0] == 'neg' and exponent[1][0] == 'num':
            exponent = ('num', -exponent[1][1])
        if exponent[0] == 'num' and float(exponent[1]).is_integer() and abs(exponent[1]) <= 64:
            bc.emit('POWI', int(exponent[1]))
        else:
            compile_ast(exponent, resolve, bc, False)
            bc.emit('POW')
    elif kind == 'call':
        for a in node[2]:
            compile_ast(a, resolve, bc, False)
        bc.emit(_CALL_OPS[node[1]])
    else:
        raise ParseError('bad node %r' % (node,))


# ---------------------------------------------------------------------------------------------
# Symbolic differentiation (Lepton semantics: d step/dx = 0, d delta/dx = 0)
# ---------------------------------------------------------------------------------------------

_ZERO, _ONE = ('num', 0.0), ('num', 1.0)


def _is(node, value):
    return node[0] == 'num' and node[1] == value


def _add(a, b):
    if _is(a, 0.0):
        return b
    if _is(b, 0.0):
        return a
    if a[0] == 'num' and b[0] == 'num':
        return ('num', a[1] + b[1])
    return ('add', a, b)


def _sub(a, b):
    if _is(b, 0.0):
        return a
    if _is(a, 0.0):
        return _neg(b)
    if a[0] == 'num' and b[0] == 'num':
        return ('num', a[1] - b[1])
    return ('sub', a, b)


def _neg(a):
    if a[0] == 'num':
        return ('num', -a[1])
    if a[0] == 'neg':
        return a[1]
    return ('neg', a)


def _mul(a, b):
    if _is(a, 0.0) or _is(b, 0.0):
        return _ZERO
    if _is(a, 1.0):
        return b
    if _is(b, 1.0):
        return a
    if a[0] == 'num' and b[0] == 'num':
        return ('num', a[1]*b[1])
    return ('mul', a, b)


def _div(a, b):
    if _is(a, 0.0):
        return _ZERO
    if _is(b, 1.0):
        return a
    return ('div', a, b)


def _pow(a, b):
    if _is(b, 0.0):
        return _ONE
    if _is(b, 1.0):
        return a
    return ('pow', a, b)


def _call(name, *args):
    return ('call', name, tuple(args))


def simplify(node):
    """Constant folding and x*1, x+0, x/1 identities (keeps coefficient expressions minimal)."""
    kind = node[0]
    if kind in ('num', 'var'):
        return node
    if kind == 'call':
        args = tuple(simplify(a) for a in node[2])
        if all(a[0] == 'num' for a in args) and node[1] in _FUNCS and node[1] != 'deriv':
            try:
                return ('num', float(_FUNCS[node[1]](*[a[1] for a in args])))
            except (ValueError, OverflowError, ZeroDivisionError):
                pass
        return ('call', node[1], args)
    parts = [simplify(a) for a in node[1:]]
    if kind == 'add':
        return _add(*parts)
    if kind == 'sub':
        return _sub(*parts)
    if kind == 'mul':
        return _mul(*parts)
    if kind == 'div':
        a, b = parts
        if a[0] == 'num' and b[0] == 'num' and b[1] != 0:
            return ('num', a[1]/b[1])
        return _div(a, b)
    if kind == 'neg':
        return _neg(parts[0])
    if kind == 'pow':
        a, b = parts
        if a[0] == 'num' and b[0] == 'num':
            try:
                return ('num', float(a[1]**b[1]))
            except (ValueError, OverflowError, ZeroDivisionError):
                pass
        return _pow(a, b)
    return (kind,) + tuple(parts)


def diff(node, x):
    """d(node)/d(x) for an inlined AST; ``x`` is a variable name."""
    kind = node[0]
    if kind == 'num':
        return _ZERO
    if kind == 'var':
        return _ONE if node[1] == x else _ZERO
    if kind == 'add':
        return _add(diff(node[1], x), diff(node[2], x))
    if kind == 'sub':
        return _sub(diff(node[1], x), diff(node[2], x))
    if kind == 'neg':
        return _neg(diff(node[1], x))
    if kind == 'mul':
        a, b = node[1], node[2]
        return _add(_mul(diff(a, x), b), _mul(a, diff(b, x)))
    if kind == 'div':
        a, b = node[1], node[2]
        da, db = diff(a, x), diff(b, x)
        if _is(db, 0.0):
            return _div(da, b)
        return _div(_sub(_mul(da, b), _mul(a, db)), _pow(b, ('num', 2.0)))
    if kind == 'pow':
        a, b = node[1], node[2]
        da, db = diff(a, x), diff(b, x)
        if b[0] == 'neg' and b[1][0] == 'num':
            b = ('num', -b[1][1])
        if _is(db, 0.0):
            if b[0] == 'num':
                return _mul(_mul(b, _pow(a, ('num', b[1] - 1.0))), da)
            return _mul(_mul(b, _pow(a, _sub(b, _ONE))), da)
        return _mul(node, _add(_mul(db, _call('log', a)), _div(_mul(b, da), a)))
    if kind == 'call':
        name, args = node[1], node[2]
        if name in ('step', 'delta', 'floor', 'ceil'):
            return _ZERO
        a = args[0]
        da = diff(a, x)
        if name == 'select':
            return _call('select', a, diff(args[1], x), diff(args[2], x))
        if name in ('min', 'max'):
            first = _call('step', _sub(args[1], a)) if name == 'min' else _call('step', _sub(a, args[1]))
            return _call('select', first, da, diff(args[1], x))
        if _is(da, 0.0):
            return _ZERO
        if name == 'sqrt':
            return _div(da, _mul(('num', 2.0), node))
        if name == 'exp':
            return _mul(node, da)
        if name == 'log':
            return _div(da, a)
        if name == 'sin':
            return _mul(_call('cos', a), da)
        if name == 'cos':
            return _neg(_mul(_call('sin', a), da))
        if name == 'tan':
            return _div(da, _pow(_call('cos', a), ('num', 2.0)))
        if name == 'erf':
            return _mul(_mul(('num', 2.0/math.sqrt(math.pi)), _call('exp', _neg(_pow(a, ('num', 2.0))))), da)
        if name == 'erfc':
            return _mul(_mul(('num', -2.0/math.sqrt(math.pi)), _call('exp', _neg(_pow(a, ('num', 2.0))))), da)
        if name == 'abs':
            return _mul(_sub(_mul(('num', 2.0), _call('step', a)), _ONE), da)
    raise ParseError('cannot differentiate %r' % (node,))

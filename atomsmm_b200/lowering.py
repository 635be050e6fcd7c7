"""
Lowering of atomsmm descriptions onto the C ABI (host logic, no physics evaluation).

* ``classify_pair_force`` / ``classify_bond_force``: decide which hand-written kernel family an
  energy string belongs to.  Recognition is *numerical*: the parsed string is evaluated on the
  host at sample points and compared with the closed form of each candidate family, which makes it
  insensitive to whitespace, ordering of auxiliary definitions, duplicated mixing rules and
  algebraic rearrangement (SURVEY hard part 4).  Anything outside the closed set raises
  ``UnsupportedDescription`` -- there is no interpreter fallback for pair potentials.
* ``lower_program``: turn a recorded CustomIntegrator step program (integrators.py:131-147 emits it)
  into the flat op list of csrc/program.h: counted while-loops unrolled, data-dependent control
  flow over globals compiled into scalar-VM code, kick / drift / scale steps mapped to dedicated
  kernels, force-buffer copies (``_f2_ <- f2``, integrators.py:134-144) folded away.
"""

import math
import random
import re

import numpy as np

from . import expr as X
from . import mm

# family ids (include/atomsmm_b200.h)
PAIR_NEAR, PAIR_DAMPED, PAIR_LJC, PAIR_LJ_VIRIAL, PAIR_SOFTCORE = 1, 2, 3, 4, 5
BOND_HARMONIC, ANGLE_HARMONIC, TORSION_PERIODIC, BOND_LJC, BOND_CUSTOM, ANGLE_CUSTOM = 1, 2, 3, 4, 5, 6
(OP_PERDOF, OP_SUM, OP_GLOBAL, OP_EVAL, OP_KICK, OP_DRIFT, OP_SCALE, OP_UPDATE_STATE, OP_ENERGY,
 OP_FUSED_INNER, OP_INVALIDATE, OP_CONSTRAIN_X, OP_CONSTRAIN_V, OP_MVV_FACTOR) = range(1, 15)
ENERGY_SLOT_DLAMBDA_VDW, ENERGY_SLOT_DLAMBDA_COUL = 64, 65
OP_WORDS = 8


class UnsupportedDescription(Exception):
    pass


# ---------------------------------------------------------------------------------------------
# closed forms (float64, python) used only to recognise energy strings
# ---------------------------------------------------------------------------------------------

def _switch(u):
    u = min(max(u, 0.0), 1.0)
    return 1 + u**3*(15*u - 6*u*u - 10)


def _ljc(r, qq, sig, eps, kc):
    s6 = (sig/r)**6
    return 4*eps*s6*(s6 - 1) + kc*qq/r


def near_closed_form(r, qq, sig, eps, variant, rs, rc, kc, sign, gate=True):
    if gate and r >= rc:
        return 0.0
    u = (r - rs)/(rc - rs) if r > rs else 0.0
    if variant == 0:
        return sign*_switch(u)*_ljc(r, qq, sig, eps, kc)
    if variant == 1:
        return sign*_switch(u)*(_ljc(r, qq, sig, eps, kc) - _ljc(rc, qq, sig, eps, kc))
    from .forces import force_switch_constants
    b, f12c, f6c, f1c = force_switch_constants(rs, rc)
    f12 = f6 = f1 = 1.0
    if u > 0:
        R = u/b + 1
        f12 += (6*b**2 - 21*b + 28)*(b**3*(R**12 - 1) - 12*b**2*u - 66*b*u**2 - 220*u**3)/462 \
            + 45*(7 - 2*b)*u**4/14 - 72*u**5/7
        f6 += (6*b**2 - 3*b + 1)*(b**3*(R**6 - 1) - 6*b**2*u - 15*b*u**2 - 20*u**3) + 45*(1 - 2*b)*u**4 - 36*u**5
        f1 += 5*(b + 1)**2*(6*b**3*R*math.log(R) - 6*b**2*u - 3*b*u**2 + u**3) + u**4*(3*u - 5*b - 10)/2
    s6, c6 = (sig/r)**6, (sig/rc)**6
    return sign*(4*eps*(f12*s6*s6 - f6*s6) + kc*qq*f1/r - (4*eps*(f12c*c6*c6 - f6c*c6) + kc*qq*f1c/rc))


def damped_closed_form(r, qq, sig, eps, alpha, rs, rc, degree, kc):
    s6 = (sig/r)**6
    core = 4*eps*s6*(s6 - 1) + math.erfc(alpha*r)*kc*qq/r
    if degree == 1:
        return core           # the built-in switch is applied by the force object, not the string
    u = (r**degree - rs**degree)/(rc**degree - rs**degree)
    return _switch(u if r > rs else 0.0)*core


def softcore_closed_form(r, qq, sig, eps, kc, lam_v, lam_c):
    x = (r/sig)**6 + 0.5*(1 - lam_v)
    return 4*lam_v*eps*(1 - x)/x**2 + kc*lam_c*qq/r


_SAMPLES = [(0.42, -0.84, 0.31, 0.25, 0.65, 0.10), (-0.3, -0.41, 0.25, 0.34, 0.07, 0.71), (0.0, 0.5, 0.2, 0.3, 0.0, 0.4)]


def _resolve(name, globals_, defs):
    if name in globals_:
        return float(globals_[name])
    if name in defs:
        try:
            return float(X.evaluate(X.substitute(defs[name], defs), globals_))
        except KeyError:
            return None
    return None


def _matches(user, closed, radii, tol=2e-7):
    worst, scale = 0.0, 1e-300
    for (q1, q2, s1, s2, e1, e2) in _SAMPLES:
        for r in radii:
            try:
                a = user(r, q1, q2, s1, s2, e1, e2)
                b = closed(r, q1*q2, 0.5*(s1 + s2), math.sqrt(e1*e2))
            except (ValueError, ZeroDivisionError, OverflowError, KeyError):
                return False
            worst = max(worst, abs(a - b))
            scale = max(scale, abs(b))
    return worst <= tol*scale


def classify_pair_force(force, parameters=None):
    """Return (family, cutoff, params, info) for a CustomNonbondedForce description."""
    names = [force.getPerParticleParameterName(k) for k in range(force.getNumPerParticleParameters())]
    if names not in (['charge', 'sigma', 'epsilon'], ['sigma', 'epsilon']):
        raise UnsupportedDescription('per-particle parameters %r: only (charge, sigma, epsilon) or (sigma, epsilon) '
                                     'without parameter offsets are supported on the hot path' % (names,))
    if force.getNonbondedMethod() != mm.CustomNonbondedForce.CutoffPeriodic:
        raise UnsupportedDescription('pair forces must use CutoffPeriodic')
    partition = None
    if force.getNumInteractionGroups() > 0:
        # the one shape atomsmm builds (systems.py:392): a set of atoms against its complement
        if force.getNumInteractionGroups() != 1:
            raise UnsupportedDescription('only a single interaction group is supported')
        set1, set2 = force.getInteractionGroupParameters(0)
        everything = set(range(force.getNumParticles()))
        if set(set1) & set(set2) or set(set1) | set(set2) != everything:
            raise UnsupportedDescription('interaction groups must be a set of atoms against its complement')
        partition = sorted(set1)
    globals_ = {force.getGlobalParameterName(k): force.getGlobalParameterDefaultValue(k)
                for k in range(force.getNumGlobalParameters())}
    if parameters:
        globals_.update({k: v for k, v in parameters.items() if k in globals_})
    main, defs = X.parse(force.getEnergyFunction())
    inlined = X.substitute(main, defs)
    cutoff = force.getCutoffDistance().value_in_md_units()
    omm_switch = force.getUseSwitchingFunction()
    rswitch = force.getSwitchingDistance().value_in_md_units()

    def user(r, q1, q2, s1, s2, e1, e2):
        env = dict(globals_, r=r, charge1=q1, charge2=q2, sigma1=s1, sigma2=s2, epsilon1=e1, epsilon2=e2)
        return X.evaluate(inlined, env)

    get = lambda name: _resolve(name, globals_, defs)
    kc = get('Kc')
    # --- near family -------------------------------------------------------------------------
    rs0, rc0 = get('rs0'), get('rc0')
    if rs0 is not None and rc0 is not None and not omm_switch:
        radii = list(np.linspace(0.12, min(rc0, cutoff)*0.9995, 23)) + [0.5*(rs0 + rc0), rs0*1.0001]
        gated = cutoff > rc0*(1 + 1e-12)
        if gated:
            radii += [rc0*1.001, 0.5*(rc0 + cutoff)]
        for variant in (0, 1, 2):
            for sign in (1.0, -1.0):
                for use_c in (1.0, 0.0):
                    k = (kc if kc is not None else 0.0)*use_c
                    if use_c and kc is None:
                        continue
                    closed = lambda r, qq, s, e: near_closed_form(r, qq, s, e, variant, rs0, rc0, k, sign, gated)
                    if _matches(user, closed, radii, 5e-6 if variant == 2 else 2e-9):
                        return PAIR_NEAR, cutoff, [float(variant), rs0, rc0, k if use_c else 0.0, sign, use_c], \
                            dict(name='near', variant=variant, sign=sign)
    # --- damped-smoothed ---------------------------------------------------------------------
    alpha, rsw, rcut = get('alpha'), get('rswitch'), get('rcut')
    if alpha is not None and kc is not None and rsw is not None and rcut is not None:
        degree = get('d')
        degree = 1 if degree is None else int(round(degree))
        if (degree == 1) == bool(omm_switch):
            radii = list(np.linspace(0.12, cutoff*0.9995, 25)) + [0.5*(rsw + rcut)]
            closed = lambda r, qq, s, e: damped_closed_form(r, qq, s, e, alpha, rsw, rcut, degree, kc)
            if _matches(user, closed, radii, 2e-9):
                if degree == 1 and (abs(rswitch - rsw) > 1e-12 or abs(cutoff - rcut) > 1e-12):
                    raise UnsupportedDescription('DampedSmoothedForce: switching distance mismatch')
                return PAIR_DAMPED, cutoff, [alpha, rsw, rcut, float(degree), kc], dict(name='damped', degree=degree)
    # --- LJ virial ---------------------------------------------------------------------------
    radii = list(np.linspace(0.12, cutoff*0.9995, 25))
    closed = lambda r, qq, s, e: 24*e*(2*(s/r)**12 - (s/r)**6)
    if _matches(user, closed, radii, 2e-9):
        return PAIR_LJ_VIRIAL, cutoff, [1.0 if omm_switch else 0.0, rswitch, cutoff], dict(name='lj-virial')
    # --- soft core ---------------------------------------------------------------------------
    for lv, lc in (('lambda_vdw', 'lambda_coul'), ('lambda', None), ('lambda_vdw', None)):
        if lv in globals_ and (lc is None or lc in globals_):
            lam_v = globals_[lv]
            lam_c = globals_[lc] if lc else 0.0
            k = kc if (kc is not None and lc) else 0.0
            closed = lambda r, qq, s, e: softcore_closed_form(r, qq, s, e, k, lam_v, lam_c)
            samples_ok = _matches(user, closed, radii, 2e-9)
            if samples_ok:
                if partition is not None and k != 0.0:
                    raise UnsupportedDescription('interaction groups are only supported without a Coulomb term')
                return PAIR_SOFTCORE, cutoff, [k, lam_v, lam_c, 1.0 if omm_switch else 0.0, rswitch, cutoff,
                                               1.0 if partition is not None else 0.0], \
                    dict(name='softcore', lambda_vdw=lv, lambda_coul=lc, partition=partition)
    if partition is not None:
        raise UnsupportedDescription('interaction groups are only supported for the soft-core family')
    # --- plain LJ + Coulomb ------------------------------------------------------------------
    if kc is not None:
        closed = lambda r, qq, s, e: _ljc(r, qq, s, e, kc)
        if _matches(user, closed, radii, 2e-9):
            return PAIR_LJC, cutoff, [kc, 1.0, 0.0, 0.0, 0.0, 1.0 if omm_switch else 0.0, rswitch, cutoff], \
                dict(name='ljc-plain', switch_all=omm_switch)
    raise UnsupportedDescription('energy function is outside the supported pair families:\n  %s'
                                 % force.getEnergyFunction())


def tail_integral(energy, rc, rs=None, npts=200):
    """int_rc^inf E(r) r^2 dr (+ int_rs^rc (1-S) E r^2 dr), Gauss-Legendre after x = rc/r."""
    nodes, weights = np.polynomial.legendre.leggauss(npts)
    x = 0.5*(nodes + 1)
    w = 0.5*weights
    total = sum(wi*energy(rc/xi)*rc**3/xi**4 for xi, wi in zip(x, w))
    if rs is not None and rs < rc:
        r = rs + (rc - rs)*x
        for ri, wi in zip(r, w):
            t = (ri - rs)/(rc - rs)
            S = 1 + t**3*(-10 + t*(15 - 6*t))
            total += (rc - rs)*wi*(1 - S)*energy(ri)*ri*ri
    return total


def long_range_correction(classes, counts, pair_energy, rc, rs, volume):
    """2 pi N^2 / V <int E r^2 dr> averaged over class pairs."""
    n = float(sum(counts))
    total = 0.0
    for a in range(len(classes)):
        for b in range(a, len(classes)):
            count = counts[a]*(counts[a] + 1)/2 if a == b else counts[a]*counts[b]
            sig = 0.5*(classes[a][0] + classes[b][0])
            eps = math.sqrt(classes[a][1]*classes[b][1])
            if eps == 0.0:
                continue
            total += count*tail_integral(lambda r: pair_energy(r, sig, eps), rc, rs, 96)
    total /= n*(n + 1)/2
    return 2*math.pi*n*n*total/volume


def classify_bond_force(force, parameters=None):
    """(family, per-term table, gparams, code) for a CustomBondForce description."""
    names = [force.getPerBondParameterName(k) for k in range(force.getNumPerBondParameters())]
    globals_ = {force.getGlobalParameterName(k): force.getGlobalParameterDefaultValue(k)
                for k in range(force.getNumGlobalParameters())}
    if parameters:
        globals_.update({k: v for k, v in parameters.items() if k in globals_})
    main, defs = X.parse(force.getEnergyFunction())
    inlined = X.substitute(main, defs)
    if names == ['chargeprod', 'sigma', 'epsilon']:
        kc = _resolve('Kc', globals_, defs)
        if kc is not None:
            ok = True
            for (qq, _, s, _, e, _) in _SAMPLES:
                for r in (0.11, 0.2, 0.33, 0.71):
                    a = X.evaluate(inlined, dict(globals_, r=r, chargeprod=qq, sigma=s, epsilon=e))
                    b = _ljc(r, qq, s, e, kc)
                    ok = ok and abs(a - b) <= 1e-10*max(1.0, abs(b))
            if ok:
                return BOND_LJC, [kc, 0.0], None
    return BOND_CUSTOM, None, compile_custom(inlined, 'r', names, globals_)


def compile_custom(inlined, variable, names, globals_):
    """Bytecode of E(s) and dE/ds for a custom bond / angle energy."""
    index = {variable: 0}
    index.update({name: 1 + k for k, name in enumerate(names)})

    def resolve(name):
        if name in index:
            return 'PUSHV', index[name]
        if name in globals_:
            return 'PUSHC', bc.const(globals_[name])
        raise UnsupportedDescription('unknown symbol %r in custom energy' % name)
    bc = X.Bytecode()
    X.compile_ast(inlined, resolve, bc)
    n_e = len(bc.code)//2
    code_e = list(bc.code)
    bc.code = []
    X.compile_ast(X.diff(inlined, variable), resolve, bc)
    return dict(code_e=code_e, n_e=n_e, code_de=list(bc.code), n_de=len(bc.code)//2, consts=list(bc.consts))


# ---------------------------------------------------------------------------------------------
# integrator programs
# ---------------------------------------------------------------------------------------------

_FORCE = re.compile(r'^f([0-9]*)$')
_RANDOM = ('gaussian', 'uniform', 'random')
CI = mm.CustomIntegrator


class _Block(object):
    def __init__(self, kind, condition=None):
        self.kind, self.condition, self.body = kind, condition, []


def _tree(steps):
    root = _Block('root')
    stack = [root]
    for kind, variable, expression in steps:
        if kind in (CI.IfBlockStart, CI.WhileBlockStart):
            block = _Block('if' if kind == CI.IfBlockStart else 'while', expression)
            stack[-1].body.append(block)
            stack.append(block)
        elif kind == CI.BlockEnd:
            if len(stack) == 1:
                raise UnsupportedDescription('unbalanced endBlock in integrator program')
            stack.pop()
        else:
            stack[-1].body.append((kind, variable, expression))
    if len(stack) != 1:
        raise UnsupportedDescription('unterminated block in integrator program')
    return root


def _assigned(body, out=None):
    out = set() if out is None else out
    for item in body:
        if isinstance(item, _Block):
            _assigned(item.body, out)
        else:
            out.add(item[1])
    return out


def _only_global(body):
    for item in body:
        if isinstance(item, _Block):
            if not _only_global(item.body):
                return False
        elif item[0] != CI.ComputeGlobal:
            return False
    return True


def _unroll(body):
    """Unroll ``c <- k0; while (c < N) { ...; c <- c + 1 }`` when the body is not purely global."""
    out = []
    for idx, item in enumerate(body):
        if not isinstance(item, _Block):
            out.append(item)
            continue
        item.body = _unroll(item.body)
        if _only_global(item.body):
            out.append(item)
            continue
        if item.kind == 'while':
            m = re.match(r'^\s*([A-Za-z_]\w*)\s*<\s*([0-9.eE+-]+)\s*$', item.condition)
            last = item.body[-1] if item.body else None
            prev = out[-1] if out else None
            if m and last and not isinstance(last, _Block) and prev and not isinstance(prev, _Block):
                counter, bound = m.group(1), float(m.group(2))
                inc = re.sub(r'\s', '', last[2])
                try:
                    start = float(prev[2])
                except ValueError:
                    start = None
                if (last[0] == CI.ComputeGlobal and last[1] == counter and inc == counter + '+1' and
                        prev[0] == CI.ComputeGlobal and prev[1] == counter and start is not None and
                        counter not in _assigned(item.body[:-1])):
                    trips = max(0, int(math.ceil(bound - start)))
                    for _ in range(trips):
                        out.extend(item.body[:-1])
                    out.append((CI.ComputeGlobal, counter, repr(start + trips)))
                    continue
        raise UnsupportedDescription('%s block with per-DOF steps and a data-dependent condition (%s) cannot be '
                                     'lowered without host round trips' % (item.kind, item.condition))
    return out


class Program(object):
    def __init__(self):
        self.ops = []
        self.bc = X.Bytecode()
        self.global_names = ['dt']
        self.global_values = [0.0]
        self.perdof_names = []
        self.force_slots = set()
        self.uses_random = False

    def gindex(self, name):
        return self.global_names.index(name)

    def new_global(self, name, value=0.0):
        self.global_names.append(name)
        self.global_values.append(float(value))
        return len(self.global_names) - 1

    def op(self, kind, a=0, b=0, c=0, d=0, e=0, f=0, g=0):
        self.ops.append([kind, a, b, c, d, e, f, g])

    def packed_ops(self):
        return np.array(self.ops, dtype=np.int32).reshape(-1, OP_WORDS)


def lower_program(integrator, group_mask_all=0xffffffff, parameters=None, fast=True, constrained=True,
                  derivative_slots=None):
    """CustomIntegrator description -> Program (ops + bytecode + initial globals)."""
    P = Program()
    P.global_values[0] = integrator._dt
    for k in range(integrator.getNumGlobalVariables()):
        P.new_global(integrator.getGlobalVariableName(k), integrator._global_values[k])
    for name, value in (parameters or {}).items():
        if name not in P.global_names:
            P.new_global(name, value)
    P.perdof_names = [integrator.getPerDofVariableName(k) for k in range(integrator.getNumPerDofVariables())]
    steps = [tuple(integrator.getComputationStep(k)) for k in range(integrator.getNumComputations())]
    if not constrained:
        # a System without constraints: constraining positions / velocities does nothing
        steps = [s for s in steps if s[0] not in (CI.ConstrainPositions, CI.ConstrainVelocities)]
    context_parameters = set(parameters or {})
    derivative_slots = dict(derivative_slots or {})
    body = _unroll(_tree(steps).body)
    body = _fold_force_copies(body, P)
    body, dead_stores = _strip_dead_constant_stores(body)
    assigned_globals = _assigned(body)

    def resolve_perdof(name):
        if name == 'x':
            return 'PUSHV', 0
        if name == 'v':
            return 'PUSHV', 1
        if name == 'm':
            return 'PUSHM', 0
        m = _FORCE.match(name)
        if m:
            slot = 32 if m.group(1) == '' else int(m.group(1))
            return 'PUSHF', slot
        if name in P.perdof_names:
            return 'PUSHV', 2 + P.perdof_names.index(name)
        if name == 'gaussian':
            P.uses_random = True
            return 'GAUSS', 0
        if name in ('uniform', 'random'):
            P.uses_random = True
            return 'UNIF', 0
        if name in P.global_names:
            return 'PUSHG', P.gindex(name)
        if name.startswith('__dE_'):
            if name[5:] not in derivative_slots:
                raise UnsupportedDescription('deriv(energy, %s): no force of the System has a derivative with '
                                             'respect to this parameter' % name[5:])
            return 'PUSHE', derivative_slots[name[5:]]
        raise UnsupportedDescription('unknown variable %r in integrator expression' % name)

    def resolve_global(name):
        if name in ('x', 'v', 'm', 'f') or _FORCE.match(name) or name in P.perdof_names:
            raise UnsupportedDescription('per-DOF variable %r used in a global expression' % name)
        return resolve_perdof(name)

    def emit_code(ast, resolver):
        start = len(P.bc.code)
        X.compile_ast(ast, resolver, P.bc)
        return start, (len(P.bc.code) - start)//2

    def force_slots_of(ast):
        slots = []
        for name in sorted(X.free_symbols(ast)):
            m = _FORCE.match(name)
            if m:
                slots.append(32 if m.group(1) == '' else int(m.group(1)))
        return slots

    def ensure_forces(ast):
        for slot in force_slots_of(ast):
            mask = group_mask_all if slot == 32 else (1 << slot)
            P.op(OP_EVAL, _as_i32(mask), slot)
            P.force_slots.add(slot)

    # ---- coefficient management for the fast kernels ------------------------------------------
    prologue = []         # (gindex, ast) evaluated once at the start of every step
    coef_cache = {}

    def coefficient(ast):
        """Global slot holding the value of a per-DOF-free AST at the time of use."""
        ast = X.simplify(ast)
        if ast[0] == 'var' and ast[1] in P.global_names:
            return P.gindex(ast[1]), None
        key = X.to_string(ast)
        symbols = X.free_symbols(ast)
        if any(s in _RANDOM for s in symbols):
            return None, None
        invariant = not (symbols & assigned_globals)
        if invariant and key in coef_cache:
            return coef_cache[key], None
        slot = P.new_global('_coef%d' % len(P.global_names))
        if invariant:
            coef_cache[key] = slot
            prologue.append((slot, ast))
            return slot, None
        return slot, ast

    def emit_global_assign(slot, ast):
        start = len(P.bc.code)
        X.compile_ast(ast, resolve_global, P.bc)
        P.bc.emit('STOREG', slot)
        P.op(OP_GLOBAL, 0, start, (len(P.bc.code) - start)//2)

    def try_fast(variable, ast):
        if not fast:
            return False
        shape = _recognise_fast(variable, ast, P)
        if shape is None:
            return False
        kind, coef_ast, extra = shape
        slot, pending = coefficient(coef_ast)
        if slot is None:
            return False
        if pending is not None:
            emit_global_assign(slot, pending)
        if kind == 'kick':
            (fa, sa), second = extra
            ensure_forces(ast)
            terms = [(fa, slot, 1 if sa > 0 else -1)]
            if second is not None:
                terms.append((second[0], slot, 1 if second[1] > 0 else -1))
            P.ops.append(['KICK', terms, -1])
        elif kind == 'drift':
            P.op(OP_DRIFT, slot)
        else:
            P.op(OP_SCALE, slot)
        return True

    def lower_global_run(items):
        """Consecutive global statements and global-only blocks -> one scalar-VM op.  A run is cut at
        statements that read deriv(energy, p) (the derivative must be evaluated first, with the
        parameters as they are at that point) and after statements that move a context parameter
        (forces depend on it)."""
        pieces, current = [], []
        for item in items:
            flat = not isinstance(item, _Block)
            reads = flat and 'deriv(' in item[2]
            writes = (item[1] in context_parameters) if flat else bool(_assigned(item.body) & context_parameters)
            if not flat and any('deriv(' in t for t in _block_texts(item)):
                raise UnsupportedDescription('deriv(energy, .) inside an if/while block is not supported')
            if reads and current:
                pieces.append((current, False, False))
                current = []
            current.append(item)
            if reads or writes:
                pieces.append((current, reads, writes))
                current = []
        if current:
            pieces.append((current, False, False))
        for piece, reads, writes in pieces:
            if reads:
                P.op(OP_ENERGY)
            _lower_global_piece(piece)
            if writes:
                P.op(OP_INVALIDATE)

    def _lower_global_piece(items):
        start = len(P.bc.code)
        base = start//2

        def here():
            return (len(P.bc.code) - start)//2

        def emit_items(seq):
            for item in seq:
                if isinstance(item, _Block):
                    lhs, opcode, rhs = X.parse_condition(item.condition)
                    top = here()
                    X.compile_ast(lhs, resolve_global, P.bc)
                    X.compile_ast(rhs, resolve_global, P.bc)
                    P.bc.emit('CMP', opcode)
                    jz = len(P.bc.code)
                    P.bc.emit('JMPZ', 0)
                    emit_items(item.body)
                    if item.kind == 'while':
                        P.bc.emit('JMP', top)
                    P.bc.code[jz + 1] = here()
                else:
                    _, variable, expression = item
                    if variable not in P.global_names:
                        raise UnsupportedDescription('assignment to unknown global %r' % variable)
                    X.compile_ast(_replace_deriv(X.parse_inlined(expression)), resolve_global, P.bc)
                    P.bc.emit('STOREG', P.gindex(variable))
        emit_items(items)
        del base
        P.op(OP_GLOBAL, 0, start, (len(P.bc.code) - start)//2)

    pending_globals = []
    serial = 0
    for item in body:
        is_global = isinstance(item, _Block) or item[0] == CI.ComputeGlobal
        if is_global:
            pending_globals.append(item)
            continue
        if pending_globals:
            lower_global_run(pending_globals)
            pending_globals = []
        kind, variable, expression = item
        if kind == CI.UpdateContextState:
            P.op(OP_UPDATE_STATE)
        elif kind == CI.ConstrainPositions:
            P.op(OP_CONSTRAIN_X)
        elif kind == CI.ConstrainVelocities:
            P.op(OP_CONSTRAIN_V)
        elif kind == CI.ComputeSum:
            ast = X.parse_inlined(expression)
            if variable not in P.global_names:
                raise UnsupportedDescription('sum target %r is not a global variable' % variable)
            ensure_forces(ast)
            is_mvv = re.sub(r'\s', '', expression) in ('m*v*v', 'm*v^2', 'v*v*m')
            start, length = emit_code(ast, resolve_perdof)
            P.op(OP_SUM, P.gindex(variable), start, length, 1 if is_mvv else 0)
        elif kind == CI.ComputePerDof:
            ast = X.parse_inlined(expression)
            if variable == 'x':
                target = 0
            elif variable == 'v':
                target = 1
            elif variable in P.perdof_names:
                target = 2 + P.perdof_names.index(variable)
            else:
                raise UnsupportedDescription('per-DOF target %r is not declared' % variable)
            if try_fast(variable, ast):
                continue
            ensure_forces(ast)
            before = P.uses_random
            P.uses_random = False
            start, length = emit_code(ast, resolve_perdof)
            rnd = P.uses_random
            P.uses_random = before or rnd
            serial += 1
            P.op(OP_PERDOF, target, start, length, 1 if rnd else 0, serial)
        else:
            raise UnsupportedDescription('unsupported step kind %d' % kind)
    if pending_globals:
        lower_global_run(pending_globals)
    for name, value in dead_stores.items():
        if name in P.global_names:
            prologue.append((P.gindex(name), ('num', value)))
    if prologue:
        start = len(P.bc.code)
        for slot, ast in prologue:
            X.compile_ast(ast, resolve_global, P.bc)
            P.bc.emit('STOREG', slot)
        # d = 1 marks the prologue: it only depends on globals the program never assigns, so the engine
        # runs it once and again only after the host changed a global
        P.ops.insert(0, [OP_GLOBAL, 0, start, (len(P.bc.code) - start)//2, 1, 0, 0, 0])
    _fuse_kicks(P)
    if fast:
        _fuse_velocity_ops(P)
        _chain_scale_blocks(P)
    return P


MAX_KICK_TERMS = 6


def _replace_deriv(node):
    """deriv(energy, p) -> pseudo variable __dE_p (resolved to a VM_PUSHE of the derivative slot)."""
    if not isinstance(node, tuple):
        return node
    if node[0] == 'call' and node[1] == 'deriv':
        target, parameter = node[2]
        if target != ('var', 'energy') or parameter[0] != 'var':
            raise UnsupportedDescription('only deriv(energy, parameter) is supported')
        return ('var', '__dE_' + parameter[1])
    if node[0] == 'call':
        return ('call', node[1], [_replace_deriv(a) for a in node[2]])
    return tuple(_replace_deriv(c) if isinstance(c, tuple) else c for c in node)


def _block_texts(block):
    out = [block.condition]
    for item in block.body:
        out += _block_texts(item) if isinstance(item, _Block) else [item[2]]
    return out


def _fuse_kicks(P):
    """Peephole over the op list: consecutive kicks (possibly separated by force evaluations, which
    do not touch velocities) become one multi-term kick, and a drift that directly follows is folded
    into the same kernel.  Term tables {slot, coefficient, sign} go into the code pool."""
    out = []
    pending = None

    def flush():
        nonlocal pending
        if pending is not None:
            terms, drift = pending
            offset = len(P.bc.code)
            for slot, coef, sign in terms:
                P.bc.code += [slot, coef, sign]
            if len(P.bc.code) % 2:
                P.bc.code.append(0)
            out.append([OP_KICK, len(terms), offset, drift, -1, -1, 0, 0])
            pending = None

    for op in P.ops:
        if op[0] == 'KICK':
            if pending is not None and (pending[1] >= 0 or len(pending[0]) + len(op[1]) > MAX_KICK_TERMS):
                flush()
            if pending is None:
                pending = [list(op[1]), -1]
            else:
                pending[0].extend(op[1])
        elif op[0] == OP_EVAL:
            # evaluating forces neither reads nor writes v; but a drift already folded into the
            # pending kick changes x, so the evaluation must stay behind it
            if pending is not None and pending[1] >= 0:
                flush()
            out.append(op)
        elif op[0] == OP_DRIFT and pending is not None and pending[1] < 0:
            pending[1] = op[1]
            flush()
        else:
            flush()
            out.append(op)
    flush()
    P.ops = out


def _fuse_velocity_ops(P):
    """Second peephole: everything that only rescales / kicks velocities and reduces ``m*v*v``
    becomes part of ONE velocity kernel (OP_KICK with optional pre-scale, kick terms, drift,
    mvv reduction and the scalar program that consumes the sum):

        SCALE s ; [EVAL|UPDATE]* ; KICK        ->  KICK(prescale=s)    (evaluations touch neither v nor globals)
        SCALE s ; SUM mvv                      ->  KICK(0 terms, prescale=s, mvv)
        KICK ; SUM mvv                         ->  KICK(mvv)
        KICK(mvv) ; GLOBAL                     ->  KICK(mvv, scalar program)

    A Nose-Hoover block (sum, scalar update, scale) thereby costs one launch instead of four."""
    out = []
    held = None          # a SCALE waiting for the next velocity op
    passed = []          # EVAL / UPDATE ops that overtook the held scale

    def bare(prescale=-1):
        return [OP_KICK, 0, 0, -1, prescale, -1, 0, 0]

    def release():
        nonlocal held, passed
        if held is not None:
            out.append(bare(held))
            held = None
        out.extend(passed)
        passed = []

    for op in P.ops:
        kind = op[0]
        if kind == OP_SCALE:
            release()
            held = op[1]
        elif kind in (OP_EVAL, OP_UPDATE_STATE):
            if held is not None:
                passed.append(op)
            else:
                out.append(op)
        elif kind == OP_KICK:
            if held is not None:
                out.extend(passed)
                passed = []
                op = list(op)
                op[4] = held
                held = None
            out.append(list(op))
        elif kind == OP_SUM and op[4] == 1:
            if held is not None and not passed:
                out.append(bare(held))
                held = None
            else:
                release()
            if out and out[-1][0] == OP_KICK and out[-1][5] < 0:
                out[-1][5] = op[1]
            else:
                k = bare()
                k[5] = op[1]
                out.append(k)
        elif kind == OP_GLOBAL:
            release()
            if out and out[-1][0] == OP_KICK and out[-1][5] >= 0 and out[-1][7] == 0:
                out[-1][6], out[-1][7] = op[2], op[3]
            else:
                out.append(op)
        else:
            release()
            out.append(op)
    release()
    P.ops = out


def _chain_scale_blocks(P):
    """Third peephole: consecutive thermostat blocks share ONE reduction.

    After ``v <- s*v`` the sum ``m*v*v`` is exactly ``s*s`` times its previous value, so in
        KICK(mvv -> X, program A) ; KICK(no terms, prescale s, mvv -> X, program B)
    the second reduction is replaced by the scalar statement ``X <- s*s*X`` and the two programs run
    back to back in the first kernel's tail.  The rescaling itself is deferred: the product of the
    pending factors is kept in a private global and applied by the next velocity kernel (as its
    pre-scale) or, failing that, by a bare scaling kernel.  A Suzuki-Yoshida chain of n Nose-Hoover
    blocks thereby costs one reduction instead of n.  (Results differ from the literal program only
    by the rounding of s*s*X versus the re-summed value, ~1e-16 relative.)"""
    JUMPS = (X.OPCODES['JMP'], X.OPCODES['JMPZ'])
    PUSHG, MUL, STOREG = X.OPCODES['PUSHG'], X.OPCODES['MUL'], X.OPCODES['STOREG']
    product = None       # global index of the pending scale product, or None
    chain_head = None
    out = []

    def words(start, length, shift):
        code = list(P.bc.code[start:start + 2*length])
        for k in range(0, len(code), 2):
            if code[k] in JUMPS:
                code[k+1] += shift
        return code

    def flush():
        nonlocal product
        if product is not None:
            out.append([OP_KICK, 0, 0, -1, product, -1, 0, 0])
            product = None

    vs = None
    for op in P.ops:
        kind = op[0]
        bare_sum = (kind == OP_KICK and op[1] == 0 and op[3] < 0 and op[4] >= 0 and op[5] >= 0)
        head = None
        if bare_sum:
            # the kernel that produced X: the last op, provided only EVAL/UPDATE ops came after it
            for prev in reversed(out):
                if prev[0] in (OP_EVAL, OP_UPDATE_STATE):
                    continue
                if prev[0] == OP_KICK and prev[5] == op[5] and (product is None or prev is chain_head):
                    head = prev
                break
        if head is not None:
            if vs is None:
                vs = P.new_global('_vscale_product', 1.0)
            scale, target = op[4], op[5]
            merged = words(head[6], head[7], 0)
            shift = len(merged)//2
            merged += [PUSHG, scale, PUSHG, scale, MUL, 0, PUSHG, target, MUL, 0, STOREG, target]
            if product is None:
                merged += [PUSHG, scale, STOREG, vs]
            else:
                merged += [PUSHG, vs, PUSHG, scale, MUL, 0, STOREG, vs]
            shift = len(merged)//2
            merged += words(op[6], op[7], shift)
            head[6] = len(P.bc.code)
            head[7] = len(merged)//2
            P.bc.code += merged
            product = vs
            chain_head = head
            continue
        if kind in (OP_EVAL, OP_UPDATE_STATE):
            out.append(op)
            continue
        if product is not None and kind == OP_KICK:
            op = list(op)
            if op[4] >= 0:
                # the op's own pre-scale joins the product (computed by the chain head's program)
                extra = [PUSHG, vs, PUSHG, op[4], MUL, 0, STOREG, vs]
                merged = words(chain_head[6], chain_head[7], 0) + extra
                chain_head[6] = len(P.bc.code)
                chain_head[7] = len(merged)//2
                P.bc.code += merged
                if op[1] == 0 and op[3] < 0 and op[5] < 0:
                    # a pure rescaling closes the chain: X already contains every factor but this last one, so
                    # after the kernel X*s*s is the sum of the new velocities -- the engine carries it to the
                    # next thermostat block (also across the step boundary) instead of summing again
                    out.append([OP_MVV_FACTOR, chain_head[5], op[4], 0, 0, 0, 0, 0])
            op[4] = vs
            product = None
            out.append(op)
            continue
        flush()
        if kind == OP_KICK and op[1] == 0 and op[3] < 0 and op[4] >= 0 and op[5] < 0:
            # unchained thermostat block: KICK(mvv -> X, program) ; v <- s v
            for prev in reversed(out):
                if prev[0] in (OP_EVAL, OP_UPDATE_STATE):
                    continue
                if prev[0] == OP_KICK and prev[5] >= 0:
                    out.append([OP_MVV_FACTOR, prev[5], op[4], 0, 0, 0, 0, 0])
                break
        out.append(op)
    flush()
    P.ops = out


def _strip_dead_constant_stores(body):
    """Loop counters of unrolled loops are assigned constants that nothing reads any more; keep
    only their final value (written once per step) instead of one tiny kernel per assignment."""
    reads = set()

    def collect(seq):
        for item in seq:
            if isinstance(item, _Block):
                lhs, _, rhs = X.parse_condition(item.condition)
                reads.update(X.free_symbols(lhs) | X.free_symbols(rhs))
                collect(item.body)
            elif item[2]:
                main, defs = X.parse(item[2])
                reads.update(X.free_symbols(X.substitute(main, defs)))
    collect(body)
    out, final = [], {}
    for item in body:
        if not isinstance(item, _Block) and item[0] == CI.ComputeGlobal and item[1] not in reads:
            try:
                final[item[1]] = float(item[2])
                continue
            except ValueError:
                pass
        out.append(item)
    return out, final


def _as_i32(mask):
    mask &= 0xffffffff
    return mask - (1 << 32) if mask & 0x80000000 else mask


def _fold_force_copies(body, P):
    """Remove ``_fK_ <- fK`` copies that only exist because OpenMM allows one force group per
    expression (integrators.py:134-144): every read of the buffer directly follows its copy."""
    buffers = {}
    for idx, item in enumerate(body):
        if isinstance(item, _Block):
            continue
        kind, variable, expression = item
        m = re.match(r'^_f([0-9]*)_$', variable)
        if kind == CI.ComputePerDof and m and re.sub(r'\s', '', expression) == 'f' + m.group(1):
            buffers.setdefault(variable, []).append(idx)
    if not buffers:
        return body
    ok = {}
    for name, copies in buffers.items():
        ok[name] = True
        for idx, item in enumerate(body):
            if isinstance(item, _Block):
                if name in _block_symbols(item):
                    ok[name] = False
                continue
            if idx in copies:
                continue
            reads = name in X.required_variables(item[1], item[2]) if item[0] in (
                CI.ComputePerDof, CI.ComputeSum, CI.ComputeGlobal) and item[2] else False
            if reads and (idx - 1) not in copies:
                ok[name] = False
            if item[1] == name:
                ok[name] = False
    out = []
    for idx, item in enumerate(body):
        if isinstance(item, _Block):
            out.append(item)
            continue
        kind, variable, expression = item
        if variable in buffers and idx in buffers[variable] and ok[variable]:
            continue
        for name in buffers:
            if ok[name] and expression:
                expression = re.sub(r'\b%s\b' % name, name[1:-1], expression)
        out.append((kind, variable, expression))
    return out


def _block_symbols(block):
    out = set()
    for item in block.body:
        if isinstance(item, _Block):
            out |= _block_symbols(item)
        elif item[2]:
            out |= set(X.required_variables(item[1], item[2]))
    return out


def _recognise_fast(variable, ast, P):
    """Detect  v + c*(±fa ± fb)/m,  x + c*v,  c*v  by numerical probing of the expression.

    Returns (kind, coefficient_ast, extra) or None.  The coefficient AST is obtained by
    substituting probe constants for the per-DOF symbols, so it only contains globals."""
    symbols = X.free_symbols(ast)
    if any(s in _RANDOM for s in symbols):
        return None
    perdof = {s for s in symbols if s in ('x', 'v', 'm') or _FORCE.match(s) or s in P.perdof_names}
    forces = sorted(s for s in perdof if _FORCE.match(s))
    others = perdof - set(forces) - {'x', 'v', 'm'}
    if others:
        return None
    rng = random.Random(12345)
    genv = {name: rng.uniform(0.3, 1.7) for name in symbols - perdof}

    def value(**kw):
        env = dict(genv)
        env.update({s: 0.0 for s in perdof})
        env['m'] = 1.0
        env.update(kw)
        try:
            return X.evaluate(ast, env)
        except (ZeroDivisionError, ValueError, OverflowError):
            return float('nan')

    def subst(**kw):
        mapping = {s: ('num', 0.0) for s in perdof}
        mapping['m'] = ('num', 1.0)
        mapping.update({k: ('num', float(v)) for k, v in kw.items()})
        return X.substitute(ast, mapping)

    def close(a, b):
        return abs(a - b) <= 1e-12*max(1.0, abs(a), abs(b))

    if variable == 'v' and 1 <= len(forces) <= 2 and 'x' not in perdof and 'm' in perdof:
        c = [value(**{f: 1.0}) for f in forces]
        if c[0] == 0 or math.isnan(c[0]):
            return None
        signs = [1.0] + [ck/c[0] for ck in c[1:]]
        if any(not (close(s, 1.0) or close(s, -1.0)) for s in signs):
            return None
        for _ in range(4):
            v0, mass = rng.uniform(-2, 2), rng.uniform(0.5, 20)
            fv = {f: rng.uniform(-3, 3) for f in forces}
            expect = v0 + c[0]*sum(s*fv[f] for s, f in zip(signs, forces))/mass
            if not close(value(v=v0, m=mass, **fv), expect):
                return None
        coef = subst(**{forces[0]: 1.0})
        slot = lambda f: 32 if f == 'f' else int(f[1:])
        first = (slot(forces[0]), 1.0)
        second = (slot(forces[1]), round(signs[1])) if len(forces) == 2 else None
        return 'kick', coef, (first, second)
    if variable == 'x' and not forces and perdof <= {'x', 'v'} and 'v' in perdof:
        c = value(v=1.0)
        for _ in range(4):
            x0, v0 = rng.uniform(-2, 2), rng.uniform(-2, 2)
            if not close(value(x=x0, v=v0), x0 + c*v0):
                return None
        return 'drift', subst(v=1.0), None
    if variable == 'v' and not forces and perdof == {'v'}:
        c = value(v=1.0)
        for _ in range(4):
            v0 = rng.uniform(-2, 2)
            if not close(value(v=v0), c*v0):
                return None
        return 'scale', subst(v=1.0), None
    return None

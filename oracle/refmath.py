"""
ORACLE (test infrastructure only -- see oracle/__init__.py).

Float64 numpy/sympy restatement of OpenMM-Reference semantics for the force objects atomsmm
builds.  Energy strings are evaluated *generically* (sympy -> numpy), not
through closed forms, so the product's expression-family recognition and hand-written kernels
are checked against the strings themselves.

Reference call sites restated: CustomNonbondedForce (forces.py:225), CustomBondForce
(forces.py:338), NonbondedForce (forces.py:153), Context.getState (utils.py:164,
computers.py:74-88).
"""

import math
import re

import numpy as np
import scipy.special
import sympy
from sympy.parsing.sympy_parser import parse_expr

ONE_4PI_EPS0 = 138.935456          # OpenMM's ONE_4PI_EPS0 (A15), kJ nm / (mol e^2)


class step(sympy.Function):
    nargs = 1

    def fdiff(self, argindex=1):
        return sympy.S.Zero


class delta(sympy.Function):
    nargs = 1

    def fdiff(self, argindex=1):
        return sympy.S.Zero


class select(sympy.Function):
    nargs = 3

    def fdiff(self, argindex=1):
        c, a, b = self.args
        if argindex == 1:
            return sympy.S.Zero
        return select(c, 1, 0) if argindex == 2 else select(c, 0, 1)


_SYMPY_FUNCS = dict(step=step, delta=delta, select=select, sqrt=sympy.sqrt, exp=sympy.exp, log=sympy.log,
                    sin=sympy.sin, cos=sympy.cos, tan=sympy.tan, erf=sympy.erf, erfc=sympy.erfc, abs=sympy.Abs,
                    min=sympy.Min, max=sympy.Max, floor=sympy.floor, ceil=sympy.ceiling)
_NUMPY_FUNCS = {'step': lambda x: np.where(np.asarray(x) < 0, 0.0, 1.0),
                'delta': lambda x: np.where(np.asarray(x) == 0, 1.0, 0.0),
                'select': lambda c, a, b: np.where(np.asarray(c) != 0, a, b),
                'erf': scipy.special.erf, 'erfc': scipy.special.erfc}
_IDENT = re.compile(r'[A-Za-z_][A-Za-z_0-9]*')


def parse_energy(text):
    """'main; a = ..; b = ..' -> one sympy expression with definitions inlined."""
    parts = [p.strip() for p in text.split(';') if p.strip()]
    main = parts[0]
    m = re.match(r'^([A-Za-z_][A-Za-z_0-9]*)\s*=(.*)$', main)
    if m:
        main = m.group(2)
    names = set(_IDENT.findall(text)) - set(_SYMPY_FUNCS)
    local = {name: sympy.Symbol(name) for name in names}
    local.update(_SYMPY_FUNCS)

    def conv(s):
        return parse_expr(s.replace('^', '**'), local_dict=local, evaluate=True)
    expression = conv(main)
    defs = {}
    for part in parts[1:]:
        name, rhs = part.split('=', 1)
        name = name.strip()
        if name not in defs:
            defs[name] = conv(rhs)
    for _ in range(len(defs) + 1):
        present = [s for s in expression.free_symbols if s.name in defs]
        if not present:
            break
        expression = expression.subs({s: defs[s.name] for s in present})
    return expression


def _lambdify(symbols, expression):
    return sympy.lambdify(symbols, expression, modules=[_NUMPY_FUNCS, 'scipy', 'numpy'])


_PAIR_CACHE = {}


def _pair_functions(text, glob, syms):
    """(expression, E(r, ...), dE/dr(r, ...)) for an energy string with globals substituted."""
    key = (text, tuple(sorted(glob.items())), tuple(s.name for s in syms))
    if key not in _PAIR_CACHE:
        expression = parse_energy(text).subs({sympy.Symbol(k): v for k, v in glob.items()})
        unknown = expression.free_symbols - set(syms)
        if unknown:
            raise ValueError('unbound symbols %s in %s' % (unknown, text))
        _PAIR_CACHE[key] = (expression, _lambdify(syms, expression),
                            _lambdify(syms, sympy.diff(expression, sympy.Symbol('r'))))
    return _PAIR_CACHE[key]


def omm_switch(r, rs, rc):
    """OpenMM built-in switching function and derivative (A4)."""
    t = np.clip((r - rs)/(rc - rs), 0.0, 1.0)
    S = 1 + t**3*(-10 + t*(15 - 6*t))
    dS = t*t*(-30 + t*(60 - 30*t))/(rc - rs)
    return S, dS


def min_image(d, box):
    return d - box*np.round(d/box)


def all_pairs(pos, box, cutoff, periodic=True):
    """(i, j, dvec, r) for all i<j with r < cutoff (minimum image in an orthorhombic box)."""
    n = len(pos)
    if n <= 4000:
        i, j = np.triu_indices(n, 1)
        d = pos[j] - pos[i]
        if periodic:
            d = min_image(d, box)
        r = np.sqrt((d*d).sum(1))
        keep = r < cutoff if cutoff is not None else np.ones(len(r), bool)
        return i[keep], j[keep], d[keep], r[keep]
    from scipy.spatial import cKDTree
    wrapped = pos - box*np.floor(pos/box) if periodic else pos
    wrapped = np.where(wrapped >= box, wrapped - box, wrapped) if periodic else wrapped
    tree = cKDTree(wrapped, boxsize=box if periodic else None)
    pairs = tree.query_pairs(cutoff, output_type='ndarray')
    i, j = pairs[:, 0], pairs[:, 1]
    d = pos[j] - pos[i]
    if periodic:
        d = min_image(d, box)
    r = np.sqrt((d*d).sum(1))
    keep = r < cutoff
    return i[keep], j[keep], d[keep], r[keep]


def _scatter(n, i, j, fvec):
    """fvec = force on j from pair (i,j); returns per-atom forces."""
    f = np.zeros((n, 3))
    np.add.at(f, j, fvec)
    np.add.at(f, i, -fvec)
    return f


class Result(object):
    def __init__(self, n):
        self.energy = 0.0
        self.forces = np.zeros((n, 3))
        self.virial = 0.0        # sum over pairs of r . F  ( = -sum r dE/dr )
        self.pairs = None

    def add_pairs(self, i, j, d, r, e, dedr):
        self.energy += float(np.sum(e))
        fvec = (-dedr/r)[:, None]*d
        self.forces += _scatter(len(self.forces), i, j, fvec)
        self.virial += float(np.sum(-dedr*r))


def _exclusion_filter(n, i, j, excl):
    if len(excl) == 0:
        return np.ones(len(i), bool)
    e = np.asarray(excl, dtype=np.int64)
    lo, hi = np.minimum(e[:, 0], e[:, 1]), np.maximum(e[:, 0], e[:, 1])
    keys = set((lo*n + hi).tolist())
    pk = np.minimum(i, j).astype(np.int64)*n + np.maximum(i, j)
    return ~np.isin(pk, np.fromiter(keys, dtype=np.int64, count=len(keys)))


# ---------------------------------------------------------------------------------------------
# individual force objects (duck-typed on the OpenMM getter API)
# ---------------------------------------------------------------------------------------------

def eval_custom_nonbonded(force, pos, box, params=None, want_pairs=False):
    n = force.getNumParticles()
    res = Result(n)
    names = [force.getPerParticleParameterName(k) for k in range(force.getNumPerParticleParameters())]
    table = np.array([force.getParticleParameters(k) for k in range(n)], dtype=float).reshape(n, len(names))
    glob = {force.getGlobalParameterName(k): force.getGlobalParameterDefaultValue(k)
            for k in range(force.getNumGlobalParameters())}
    if params:
        glob.update({k: v for k, v in params.items() if k in glob})
    r_sym = sympy.Symbol('r')
    syms = [r_sym] + [sympy.Symbol(nm + '1') for nm in names] + [sympy.Symbol(nm + '2') for nm in names]
    expression, fe, fd = _pair_functions(force.getEnergyFunction(), glob, syms)
    method = force.getNonbondedMethod()
    periodic = method == 2
    cutoff = None if method == 0 else float(force.getCutoffDistance().value_in_md_units())
    i, j, d, r = all_pairs(pos, box, cutoff, periodic)
    excl = [force.getExclusionParticles(k) for k in range(force.getNumExclusions())]
    keep = _exclusion_filter(n, i, j, excl)
    # interaction groups (OpenMM CustomNonbondedForce.addInteractionGroup): a pair interacts iff one
    # atom is in set1 and the other in set2 of some group (reference call site: systems.py:392)
    ngroups = force.getNumInteractionGroups() if hasattr(force, 'getNumInteractionGroups') else 0
    if ngroups:
        allowed = np.zeros(len(i), bool)
        for g in range(ngroups):
            set1, set2 = force.getInteractionGroupParameters(g)
            in1 = np.zeros(n, bool); in1[list(set1)] = True
            in2 = np.zeros(n, bool); in2[list(set2)] = True
            allowed |= (in1[i] & in2[j]) | (in2[i] & in1[j])
        keep &= allowed
    i, j, d, r = i[keep], j[keep], d[keep], r[keep]
    args = [r] + [table[i, k] for k in range(len(names))] + [table[j, k] for k in range(len(names))]
    e = np.broadcast_to(fe(*args), r.shape).astype(float)
    de = np.broadcast_to(fd(*args), r.shape).astype(float)
    if force.getUseSwitchingFunction() and cutoff is not None:
        S, dS = omm_switch(r, float(force.getSwitchingDistance().value_in_md_units()), cutoff)
        e, de = S*e, S*de + e*dS
    res.add_pairs(i, j, d, r, e, de)
    if want_pairs:
        res.pairs = (i, j, r)
    if force.getUseLongRangeCorrection() and cutoff is not None and periodic:
        res.energy += custom_lrc(force, expression, syms, names, table, cutoff, box)
    return res


def custom_lrc(force, expression, syms, names, table, cutoff, box):
    """Long-range correction of a CustomNonbondedForce (A10), by numerical quadrature."""
    from scipy.integrate import quad
    n = len(table)
    classes, counts = np.unique(table, axis=0, return_counts=True)
    fe = _lambdify(syms, expression)
    use_switch = force.getUseSwitchingFunction()
    rs = float(force.getSwitchingDistance().value_in_md_units()) if use_switch else cutoff
    total = 0.0
    for a in range(len(classes)):
        for b in range(a, len(classes)):
            count = counts[a]*(counts[a] + 1)/2 if a == b else counts[a]*counts[b]
            pa, pb = list(classes[a]), list(classes[b])

            def tail(r):
                return float(fe(r, *pa, *pb))*r*r
            integral = quad(tail, cutoff, np.inf, epsabs=0, epsrel=1e-12, limit=500)[0]
            if use_switch:
                def inner(r):
                    S, _ = omm_switch(np.array(r), rs, cutoff)
                    return (1 - float(S))*float(fe(r, *pa, *pb))*r*r
                integral += quad(inner, rs, cutoff, epsabs=0, epsrel=1e-12)[0]
            total += count*integral
    total /= n*(n + 1)/2
    return 2*math.pi*n*n*total/float(np.prod(box))


def _eval_pairlist_expression(energy_text, names, glob, idx_i, idx_j, table, pos, box, periodic, n):
    res = Result(n)
    if len(idx_i) == 0:
        return res
    r_sym = sympy.Symbol('r')
    syms = [r_sym] + [sympy.Symbol(nm) for nm in names]
    expression, fe, fd = _pair_functions(energy_text, glob, syms)
    d = pos[idx_j] - pos[idx_i]
    if periodic:
        d = min_image(d, box)
    r = np.sqrt((d*d).sum(1))
    args = [r] + [table[:, k] for k in range(len(names))]
    e = np.broadcast_to(fe(*args), r.shape).astype(float)
    de = np.broadcast_to(fd(*args), r.shape).astype(float)
    res.add_pairs(idx_i, idx_j, d, r, e, de)
    return res


def eval_custom_bond(force, pos, box, params=None):
    nb = force.getNumBonds()
    names = [force.getPerBondParameterName(k) for k in range(force.getNumPerBondParameters())]
    glob = {force.getGlobalParameterName(k): force.getGlobalParameterDefaultValue(k)
            for k in range(force.getNumGlobalParameters())}
    if params:
        glob.update({k: v for k, v in params.items() if k in glob})
    bonds = [force.getBondParameters(k) for k in range(nb)]
    ii = np.array([b[0] for b in bonds], dtype=int)
    jj = np.array([b[1] for b in bonds], dtype=int)
    table = np.array([b[2] for b in bonds], dtype=float).reshape(nb, len(names))
    return _eval_pairlist_expression(force.getEnergyFunction(), names, glob, ii, jj, table, pos, box,
                                     force.usesPeriodicBoundaryConditions(), len(pos))


def eval_harmonic_bond(force, pos, box):
    n = len(pos)
    res = Result(n)
    nb = force.getNumBonds()
    if nb == 0:
        return res
    b = [force.getBondParameters(k) for k in range(nb)]
    ii = np.array([x[0] for x in b]); jj = np.array([x[1] for x in b])
    r0 = np.array([x[2].value_in_md_units() for x in b]); k = np.array([x[3].value_in_md_units() for x in b])
    d = pos[jj] - pos[ii]
    r = np.sqrt((d*d).sum(1))
    res.add_pairs(ii, jj, d, r, 0.5*k*(r - r0)**2, k*(r - r0))
    return res


def _angle_terms(pos, ii, jj, kk):
    a = pos[ii] - pos[jj]
    b = pos[kk] - pos[jj]
    ra = np.sqrt((a*a).sum(1)); rb = np.sqrt((b*b).sum(1))
    cos = np.clip((a*b).sum(1)/(ra*rb), -1.0, 1.0)
    theta = np.arccos(cos)
    sin = np.sqrt(np.maximum(1 - cos*cos, 1e-30))
    # d theta / d r_i and d theta / d r_k
    dti = -(b/(ra*rb)[:, None] - (cos/(ra*ra))[:, None]*a)/sin[:, None]
    dtk = -(a/(ra*rb)[:, None] - (cos/(rb*rb))[:, None]*b)/sin[:, None]
    return theta, dti, dtk


def _apply_angle(res, ii, jj, kk, e, dedt, dti, dtk):
    res.energy += float(np.sum(e))
    fi = -dedt[:, None]*dti
    fk = -dedt[:, None]*dtk
    np.add.at(res.forces, ii, fi)
    np.add.at(res.forces, kk, fk)
    np.add.at(res.forces, jj, -(fi + fk))


def eval_harmonic_angle(force, pos, box):
    res = Result(len(pos))
    na = force.getNumAngles()
    if na == 0:
        return res
    a = [force.getAngleParameters(k) for k in range(na)]
    ii = np.array([x[0] for x in a]); jj = np.array([x[1] for x in a]); kk = np.array([x[2] for x in a])
    t0 = np.array([x[3].value_in_md_units() for x in a]); k = np.array([x[4].value_in_md_units() for x in a])
    theta, dti, dtk = _angle_terms(pos, ii, jj, kk)
    _apply_angle(res, ii, jj, kk, 0.5*k*(theta - t0)**2, k*(theta - t0), dti, dtk)
    return res


def eval_custom_angle(force, pos, box, params=None):
    res = Result(len(pos))
    na = force.getNumAngles()
    if na == 0:
        return res
    names = [force.getPerAngleParameterName(k) for k in range(force.getNumPerAngleParameters())]
    glob = {force.getGlobalParameterName(k): force.getGlobalParameterDefaultValue(k)
            for k in range(force.getNumGlobalParameters())}
    a = [force.getAngleParameters(k) for k in range(na)]
    ii = np.array([x[0] for x in a]); jj = np.array([x[1] for x in a]); kk = np.array([x[2] for x in a])
    table = np.array([x[3] for x in a], dtype=float).reshape(na, len(names))
    expression = parse_energy(force.getEnergyFunction()).subs({sympy.Symbol(k): v for k, v in glob.items()})
    t_sym = sympy.Symbol('theta')
    syms = [t_sym] + [sympy.Symbol(nm) for nm in names]
    fe = _lambdify(syms, expression)
    fd = _lambdify(syms, sympy.diff(expression, t_sym))
    theta, dti, dtk = _angle_terms(pos, ii, jj, kk)
    args = [theta] + [table[:, k] for k in range(len(names))]
    e = np.broadcast_to(fe(*args), theta.shape).astype(float)
    de = np.broadcast_to(fd(*args), theta.shape).astype(float)
    _apply_angle(res, ii, jj, kk, e, de, dti, dtk)
    return res


def eval_periodic_torsion(force, pos, box):
    res = Result(len(pos))
    nt = force.getNumTorsions()
    if nt == 0:
        return res
    t = [force.getTorsionParameters(k) for k in range(nt)]
    a1 = np.array([x[0] for x in t]); a2 = np.array([x[1] for x in t])
    a3 = np.array([x[2] for x in t]); a4 = np.array([x[3] for x in t])
    per = np.array([x[4] for x in t], dtype=float)
    phase = np.array([x[5].value_in_md_units() for x in t]); k = np.array([x[6].value_in_md_units() for x in t])
    # Blondel & Karplus formulation
    F = pos[a1] - pos[a2]
    G = pos[a2] - pos[a3]
    H = pos[a4] - pos[a3]
    A = np.cross(F, G)
    B = np.cross(H, G)
    gn = np.sqrt((G*G).sum(1))
    cosphi = (A*B).sum(1)/np.sqrt((A*A).sum(1)*(B*B).sum(1))
    sinphi = (np.cross(B, A)*G).sum(1)/(np.sqrt((A*A).sum(1)*(B*B).sum(1))*gn)
    phi = np.arctan2(sinphi, cosphi)
    e = k*(1 + np.cos(per*phi - phase))
    dedphi = -k*per*np.sin(per*phi - phase)
    A2 = (A*A).sum(1); B2 = (B*B).sum(1)
    fg = (F*G).sum(1); hg = (H*G).sum(1)
    dphi1 = -(gn/A2)[:, None]*A
    dphi4 = (gn/B2)[:, None]*B
    dphi2 = (gn/A2)[:, None]*A + (fg/(A2*gn))[:, None]*A - (hg/(B2*gn))[:, None]*B
    dphi3 = -(gn/B2)[:, None]*B - (fg/(A2*gn))[:, None]*A + (hg/(B2*gn))[:, None]*B
    res.energy = float(np.sum(e))
    for idx, dphi in ((a1, dphi1), (a2, dphi2), (a3, dphi3), (a4, dphi4)):
        np.add.at(res.forces, idx, -dedphi[:, None]*dphi)
    return res


# ---------------------------------------------------------------------------------------------
# NonbondedForce (NoCutoff / reaction field / PME), A8-A11, A15
# ---------------------------------------------------------------------------------------------

def pme_parameters(force, box):
    alpha, nx, ny, nz = force.getPMEParameters()
    alpha = float(alpha.value_in_md_units()) if hasattr(alpha, 'value_in_md_units') else float(alpha)
    rc = float(force.getCutoffDistance().value_in_md_units())
    if alpha == 0.0:
        tol = force.getEwaldErrorTolerance()
        alpha = math.sqrt(-math.log(2*tol))/rc
        grid = [max(6, int(math.ceil(2*alpha*L/(3*tol**0.2)))) for L in box]
    else:
        grid = [nx, ny, nz]
    return alpha, grid


def _bspline(order, w):
    """B-spline weights theta_k(w), k = 0..order-1, and derivatives; w = fractional part [n]."""
    n = len(w)
    data = np.zeros((order, n))
    ddata = np.zeros((order, n))
    data[order - 1] = 0
    data[1] = w
    data[0] = 1 - w
    for j in range(3, order):
        div = 1.0/(j - 1)
        data[j - 1] = div*w*data[j - 2]
        for k in range(1, j - 1):
            data[j - k - 1] = div*((w + k)*data[j - k - 2] + (j - k - w)*data[j - k - 1])
        data[0] = div*(1 - w)*data[0]
    ddata[0] = -data[0]
    for j in range(1, order):
        ddata[j] = data[j - 1] - data[j]
    div = 1.0/(order - 1)
    data[order - 1] = div*w*data[order - 2]
    for k in range(1, order - 1):
        data[order - k - 1] = div*((w + k)*data[order - k - 2] + (order - k - w)*data[order - k - 1])
    data[0] = div*(1 - w)*data[0]
    return data, ddata


def pme_reciprocal(pos, q, box, alpha, grid, order=5, kc=ONE_4PI_EPS0):
    """Smooth PME reciprocal energy and forces (Essmann et al. 1995), self term excluded."""
    n = len(pos)
    K = np.array(grid)
    u = (pos/box - np.floor(pos/box))*K            # scaled fractional coordinates [n,3]
    base = np.floor(u).astype(int)
    w = u - base
    th, dth = [], []
    for dim in range(3):
        t, dt = _bspline(order, w[:, dim])
        th.append(t); dth.append(dt)
    Q = np.zeros(K)
    idx = [(base[:, dim][None, :] + np.arange(order)[:, None]) % K[dim] for dim in range(3)]   # [order,n]
    for a in range(order):
        for b in range(order):
            for c in range(order):
                np.add.at(Q, (idx[0][a], idx[1][b], idx[2][c]), q*th[0][a]*th[1][b]*th[2][c])
    FQ = np.fft.fftn(Q)
    # B-spline moduli
    mods = []
    for dim in range(3):
        kk = K[dim]
        t0, _ = _bspline(order, np.zeros(1))
        bs = np.zeros(kk)
        bs[:order] = t0[:, 0]
        m = np.arange(kk)
        arg = 2*np.pi*np.outer(m, np.arange(kk))/kk
        sc = (bs[None, :]*np.cos(arg)).sum(1)
        ss = (bs[None, :]*np.sin(arg)).sum(1)
        mod = sc*sc + ss*ss
        for i in range(kk):
            if mod[i] < 1e-7:
                mod[i] = 0.5*(mod[i - 1] + mod[(i + 1) % kk])
        mods.append(mod)
    mx = np.fft.fftfreq(K[0], 1.0/K[0])/box[0]
    my = np.fft.fftfreq(K[1], 1.0/K[1])/box[1]
    mz = np.fft.fftfreq(K[2], 1.0/K[2])/box[2]
    m2 = mx[:, None, None]**2 + my[None, :, None]**2 + mz[None, None, :]**2
    V = float(np.prod(box))
    denom = np.pi*V*mods[0][:, None, None]*mods[1][None, :, None]*mods[2][None, None, :]*m2
    with np.errstate(divide='ignore', invalid='ignore'):
        eterm = kc*np.exp(-np.pi**2*m2/alpha**2)/denom
    eterm[0, 0, 0] = 0.0
    energy = 0.5*float(np.sum(eterm*(FQ.real**2 + FQ.imag**2)))
    # forces: convolve and interpolate gradient
    conv = np.fft.ifftn(eterm*FQ).real*np.prod(K)
    forces = np.zeros((n, 3))
    for a in range(order):
        for b in range(order):
            for c in range(order):
                g = conv[idx[0][a], idx[1][b], idx[2][c]]
                forces[:, 0] -= q*g*dth[0][a]*th[1][b]*th[2][c]*K[0]/box[0]
                forces[:, 1] -= q*g*th[0][a]*dth[1][b]*th[2][c]*K[1]/box[1]
                forces[:, 2] -= q*g*th[0][a]*th[1][b]*dth[2][c]*K[2]/box[2]
    return energy, forces


def nonbonded_lrc(force, box):
    n = force.getNumParticles()
    rows = getattr(force, '_particles', None)
    if rows is not None and n > 100000:
        # this repository's description classes keep MD-unit floats: read the table in bulk
        from atomsmm_b200 import mm
        table = mm.value_columns(rows, 0).reshape(n, 3)[:, 1:]
    else:
        table = np.array([[p.value_in_md_units() for p in force.getParticleParameters(k)[1:]] for k in range(n)])
    classes, counts = np.unique(table, axis=0, return_counts=True)
    rc = float(force.getCutoffDistance().value_in_md_units())
    use_switch = force.getUseSwitchingFunction()
    rs = float(force.getSwitchingDistance().value_in_md_units()) if use_switch else rc
    from scipy.integrate import quad
    s1 = s2 = s3 = 0.0
    for a in range(len(classes)):
        for b in range(a, len(classes)):
            count = counts[a]*(counts[a] + 1)/2 if a == b else counts[a]*counts[b]
            sigma = 0.5*(classes[a][0] + classes[b][0])
            eps = math.sqrt(classes[a][1]*classes[b][1])
            s1 += count*eps*sigma**12
            s2 += count*eps*sigma**6
            if use_switch and eps != 0:
                def inner(r):
                    S, _ = omm_switch(np.array(r), rs, rc)
                    return (1 - float(S))*((sigma/r)**12 - (sigma/r)**6)*r*r
                s3 += count*eps*quad(inner, rs, rc, epsabs=0, epsrel=1e-13)[0]
    norm = n*(n + 1)/2
    return 8*n*n*math.pi*(s1/norm/(9*rc**9) - s2/norm/(3*rc**3) + s3/norm)/float(np.prod(box))


def eval_nonbonded(force, pos, box, part='all'):
    """NonbondedForce.  ``part``: 'direct', 'reciprocal' or 'all' (force-group split, A9)."""
    n = force.getNumParticles()
    res = Result(n)
    method = force.getNonbondedMethod()
    prm = np.array([[p.value_in_md_units() for p in force.getParticleParameters(k)] for k in range(n)])
    q, sig, eps = prm[:, 0], prm[:, 1], prm[:, 2]
    exc = [force.getExceptionParameters(k) for k in range(force.getNumExceptions())]
    exc_i = np.array([e[0] for e in exc], dtype=int)
    exc_j = np.array([e[1] for e in exc], dtype=int)
    exc_p = np.array([[e[2].value_in_md_units(), e[3].value_in_md_units(), e[4].value_in_md_units()] for e in exc]).reshape(-1, 3)
    periodic = method >= 2
    ewald = method >= 3
    cutoff = None if method == 0 else float(force.getCutoffDistance().value_in_md_units())
    kc = ONE_4PI_EPS0
    if ewald:
        alpha, grid = pme_parameters(force, box)
    if part in ('direct', 'all'):
        i, j, d, r = all_pairs(pos, box, cutoff, periodic)
        keep = _exclusion_filter(n, i, j, list(zip(exc_i, exc_j)))
        i, j, d, r = i[keep], j[keep], d[keep], r[keep]
        s = 0.5*(sig[i] + sig[j])
        e4 = 4*np.sqrt(eps[i]*eps[j])
        x6 = (s/r)**6
        elj = e4*x6*(x6 - 1)
        dlj = -e4*(12*x6*x6 - 6*x6)/r
        if force.getUseSwitchingFunction() and cutoff is not None:
            S, dS = omm_switch(r, float(force.getSwitchingDistance().value_in_md_units()), cutoff)
            elj, dlj = S*elj, S*dlj + elj*dS
        qq = kc*q[i]*q[j]
        if ewald:
            ec = qq*scipy.special.erfc(alpha*r)/r
            dc = -qq*(scipy.special.erfc(alpha*r)/r**2 + 2*alpha/math.sqrt(math.pi)*np.exp(-(alpha*r)**2)/r)
        elif method in (1, 2):
            es = force.getReactionFieldDielectric()
            krf = (es - 1)/((2*es + 1)*cutoff**3)
            crf = 3*es/((2*es + 1)*cutoff)
            ec = qq*(1/r + krf*r*r - crf)
            dc = qq*(-1/r**2 + 2*krf*r)
        else:
            ec = qq/r
            dc = -qq/r**2
        res.add_pairs(i, j, d, r, elj + ec, dlj + dc)
        # exceptions: bare Coulomb + own LJ, no cutoff, no switch (A15)
        if len(exc) > 0:
            d = pos[exc_j] - pos[exc_i]
            if periodic:
                d = min_image(d, box)
            r = np.sqrt((d*d).sum(1))
            x6 = (exc_p[:, 1]/r)**6
            e = 4*exc_p[:, 2]*x6*(x6 - 1) + kc*exc_p[:, 0]/r
            de = -4*exc_p[:, 2]*(12*x6*x6 - 6*x6)/r - kc*exc_p[:, 0]/r**2
            if ewald:
                qq = kc*q[exc_i]*q[exc_j]
                e = e - qq*scipy.special.erf(alpha*r)/r
                de = de - qq*(2*alpha/math.sqrt(math.pi)*np.exp(-(alpha*r)**2)/r - scipy.special.erf(alpha*r)/r**2)
            res.add_pairs(exc_i, exc_j, d, r, e, de)
        if force.getUseDispersionCorrection() and periodic:
            res.energy += nonbonded_lrc(force, box)
    if ewald and part in ('reciprocal', 'all'):
        e, f = pme_reciprocal(pos, q, box, alpha, grid, 5, kc)
        res.energy += e - kc*alpha/math.sqrt(math.pi)*float(np.sum(q*q))
        res.forces += f
    return res


# ---------------------------------------------------------------------------------------------
# whole systems
# ---------------------------------------------------------------------------------------------

def _kind(force):
    for cls in type(force).__mro__:
        if cls.__name__ in ('NonbondedForce', 'CustomNonbondedForce', 'CustomBondForce', 'CustomAngleForce',
                            'HarmonicBondForce', 'HarmonicAngleForce', 'PeriodicTorsionForce', 'CMMotionRemover',
                            'MonteCarloBarostat'):
            return cls.__name__
    raise TypeError('unsupported force %r' % force)


def system_box(system):
    vec = system.getDefaultPeriodicBoxVectors()
    return np.array([float(vec[k].value_in_md_units()[k]) for k in range(3)])


def evaluate_system(system, pos, box=None, groups=None, params=None):
    """Energy (kJ/mol) and forces (kJ/mol/nm) of the forces whose group is in ``groups``.

    A NonbondedForce's reciprocal part belongs to its reciprocal-space group (if >= 0)."""
    pos = np.asarray(pos, dtype=np.float64)
    box = system_box(system) if box is None else np.asarray(box, dtype=np.float64)
    n = system.getNumParticles()
    total = Result(n)
    per_force = []
    for force in system.getForces():
        kind = _kind(force)
        g = force.getForceGroup()
        want = groups is None or g in groups
        res = None
        if kind == 'NonbondedForce':
            rg = force.getReciprocalSpaceForceGroup()
            rg = g if rg < 0 else rg
            want_r = groups is None or rg in groups
            part = 'all' if (want and want_r) else ('direct' if want else ('reciprocal' if want_r else None))
            if part:
                res = eval_nonbonded(force, pos, box, part)
        elif not want or kind in ('CMMotionRemover', 'MonteCarloBarostat'):
            res = None
        elif kind == 'CustomNonbondedForce':
            res = eval_custom_nonbonded(force, pos, box, params)
        elif kind == 'CustomBondForce':
            res = eval_custom_bond(force, pos, box, params)
        elif kind == 'CustomAngleForce':
            res = eval_custom_angle(force, pos, box, params)
        elif kind == 'HarmonicBondForce':
            res = eval_harmonic_bond(force, pos, box)
        elif kind == 'HarmonicAngleForce':
            res = eval_harmonic_angle(force, pos, box)
        elif kind == 'PeriodicTorsionForce':
            res = eval_periodic_torsion(force, pos, box)
        per_force.append(res)
        if res is not None:
            total.energy += res.energy
            total.forces += res.forces
            total.virial += res.virial
    total.per_force = per_force
    return total

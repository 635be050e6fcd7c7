"""
ORACLE (test infrastructure only -- see oracle/__init__.py).

Float64 interpreter for CustomIntegrator step programs, restating OpenMM's semantics:
per-DOF variables x v f f0..f31 m plus user ones, globals (dt first), `gaussian`/`uniform` fresh per
DOF component in per-DOF steps and once in global steps, massless particles not updated, forces of
a group re-evaluated whenever positions changed since that group was last evaluated, block
conditions `lhs op rhs`.  Reference call sites: integrators.py:113-163 (program emission and
stepping), propagators.py (the programs themselves).
"""

import re

import numpy as np
import sympy

from . import refmath

_COND = re.compile(r'^(.*?)(<=|>=|!=|<|>|=)(.*)$')
_FORCE = re.compile(r'^f([0-9]*)$')


def shake(constraints, mass, x, reference, tolerance=1e-13, sweeps=100000):
    """Position constraints: displacements along the constraint vectors of ``reference`` (the last
    constrained configuration) that restore every distance -- the equations SETTLE / SHAKE / CCMA solve
    in OpenMM.  Plain Gauss-Seidel to near machine precision."""
    x = x.copy()
    w = np.where(mass > 0, 1.0/np.where(mass > 0, mass, 1.0), 0.0)
    for _ in range(sweeps):
        worst = 0.0
        for i, j, d in constraints:
            s = x[i] - x[j]
            diff = d*d - s.dot(s)
            worst = max(worst, abs(diff)/(d*d))
            if abs(diff) > 2*tolerance*d*d:
                r = reference[i] - reference[j]
                g = diff/(2*s.dot(r)*(w[i] + w[j]))
                x[i] += g*w[i]*r
                x[j] -= g*w[j]*r
        if worst <= 2*tolerance:
            return x
    raise RuntimeError('SHAKE did not converge')


def rattle(constraints, mass, x, v, tolerance=1e-13, sweeps=100000):
    """Velocity constraints: remove the relative velocity along every constraint."""
    v = v.copy()
    w = np.where(mass > 0, 1.0/np.where(mass > 0, mass, 1.0), 0.0)
    for _ in range(sweeps):
        worst = 0.0
        for i, j, d in constraints:
            r = x[i] - x[j]
            rv = r.dot(v[i] - v[j])
            worst = max(worst, abs(rv)/(d*d))
            if abs(rv) > tolerance*d*d:
                g = rv/(r.dot(r)*(w[i] + w[j]))
                v[i] -= g*w[i]*r
                v[j] += g*w[j]*r
        if worst <= tolerance:
            return v
    raise RuntimeError('RATTLE did not converge')


class Interpreter(object):
    def __init__(self, system, integrator, positions, velocities=None, seed=0, parameters=None):
        self.system = system
        self.n = system.getNumParticles()
        self.box = refmath.system_box(system)
        self.mass = np.array([system.getParticleMass(i).value_in_md_units() for i in range(self.n)])[:, None]
        self.x = np.array(positions, dtype=np.float64).copy()
        self.v = np.zeros_like(self.x) if velocities is None else np.array(velocities, dtype=np.float64).copy()
        self.globals = {'dt': integrator._dt}
        for k in range(integrator.getNumGlobalVariables()):
            self.globals[integrator.getGlobalVariableName(k)] = float(integrator._global_values[k])
        self.parameters = dict(parameters or {})
        self.perdof = {}
        for k in range(integrator.getNumPerDofVariables()):
            value = integrator._perdof_values[k]
            self.perdof[integrator.getPerDofVariableName(k)] = \
                np.full((self.n, 3), float(value)) if np.isscalar(value) else np.array(value, dtype=np.float64)
        self.steps = [tuple(integrator.getComputationStep(k)) for k in range(integrator.getNumComputations())]
        self.rng = np.random.default_rng(seed)
        # openmm.MonteCarloBarostat forces act where the program calls UpdateContextState
        self.barostats = [dict(force=f, steps=0, counter=0, scale=0.0, attempts=0, accepted=0, log=[])
                          for f in system.getForces() if type(f).__name__ == 'MonteCarloBarostat']
        self._molecules = None
        self._force_cache = {}
        self._version = 0
        self._compiled = {}
        self.force_evaluations = 0
        # block structure: matching end for every block start
        self._end = {}
        stack = []
        for k, (kind, _, _) in enumerate(self.steps):
            if kind in (6, 7):
                stack.append(k)
            elif kind == 8:
                self._end[stack.pop()] = k

    # ---------------------------------------------------------------------------------------------
    def forces(self, group):
        key = 'all' if group is None else group
        cached = self._force_cache.get(key)
        if cached is not None and cached[0] == self._version:
            return cached[1]
        groups = None if group is None else {group}
        f = refmath.evaluate_system(self.system, self.x, self.box, groups, self.parameters).forces
        self.force_evaluations += 1
        self._force_cache[key] = (self._version, f)
        return f

    def constraints(self):
        if not hasattr(self, '_constraints'):
            self._constraints = [self.system.getConstraintParameters(k) for k in range(self.system.getNumConstraints())]
            self._constraints = [(int(i), int(j), float(getattr(d, 'value_in_md_units', lambda: d)()))
                                 for i, j, d in self._constraints]
        return self._constraints

    def potential_energy(self, groups=None):
        return refmath.evaluate_system(self.system, self.x, self.box, groups, self.parameters).energy

    def kinetic_energy(self):
        return 0.5*float(np.sum(self.mass*self.v*self.v))

    def energy_derivative(self, name, h=1e-6):
        """deriv(energy, name): central difference of the float64 potential energy (the energy is a
        smooth function of the context parameters atomsmm differentiates; error ~1e-9 relative)."""
        key = ('deriv', name, self._version, self.parameters.get(name))
        if key not in self._force_cache:
            base = dict(self.parameters)
            value = float(base[name])
            up = refmath.evaluate_system(self.system, self.x, self.box, None, dict(base, **{name: value + h})).energy
            dn = refmath.evaluate_system(self.system, self.x, self.box, None, dict(base, **{name: value - h})).energy
            self._force_cache[key] = (up - dn)/(2*h)
        return self._force_cache[key]

    def _compile(self, text):
        if text not in self._compiled:
            # deriv(energy, p) -> a pseudo variable resolved by energy_derivative
            text_in = text
            text = re.sub(r'deriv\(\s*energy\s*,\s*([A-Za-z_]\w*)\s*\)', r'__dE_\1', text)
            expression = refmath.parse_energy(text)
            symbols = sorted(expression.free_symbols, key=lambda s: s.name)
            self._compiled[text_in] = ([s.name for s in symbols], refmath._lambdify(symbols, expression))
            return self._compiled[text_in]
        return self._compiled[text]

    def _value(self, name, per_dof):
        if name == 'x':
            return self.x
        if name == 'v':
            return self.v
        if name == 'm':
            return self.mass
        m = _FORCE.match(name)
        if m and per_dof:
            return self.forces(None if m.group(1) == '' else int(m.group(1)))
        if name in self.perdof and per_dof:
            return self.perdof[name]
        if name == 'gaussian':
            return self.rng.standard_normal((self.n, 3)) if per_dof else float(self.rng.standard_normal())
        if name in ('uniform', 'random'):
            return self.rng.random((self.n, 3)) if per_dof else float(self.rng.random())
        if name in self.globals:
            return self.globals[name]
        if name in self.parameters:
            return self.parameters[name]
        if name.startswith('__dE_'):
            return self.energy_derivative(name[5:])
        raise KeyError('unknown variable %r' % name)

    def _evaluate(self, text, per_dof):
        names, function = self._compile(text)
        return function(*[self._value(name, per_dof) for name in names])

    def _condition(self, text):
        depth = 0
        for i, ch in enumerate(text):
            depth += ch == '('
            depth -= ch == ')'
            if depth == 0 and ch in '<>=!':
                op = text[i:i+2] if text[i:i+2] in ('<=', '>=', '!=') else ch
                lhs = float(self._evaluate(text[:i], False))
                rhs = float(self._evaluate(text[i+len(op):], False))
                return {'=': lhs == rhs, '<': lhs < rhs, '>': lhs > rhs, '!=': lhs != rhs,
                        '<=': lhs <= rhs, '>=': lhs >= rhs}[op]
        raise ValueError('bad condition %r' % text)

    # -- Monte Carlo barostat (OpenMM MonteCarloBarostatImpl::updateContextState, restated) -----------------
    @staticmethod
    def barostat_uniform(seed, counter):
        """SplitMix64(seed, counter) -> [0, 1): the stream of csrc/barostat.cu (b2_barostat_uniform)."""
        mask = (1 << 64) - 1
        z = (seed + 0x9e3779b97f4a7c15*(counter + 1)) & mask
        z = ((z ^ (z >> 30))*0xbf58476d1ce4e5b9) & mask
        z = ((z ^ (z >> 27))*0x94d049bb133111eb) & mask
        z ^= z >> 31
        return (z >> 11)/9007199254740992.0

    def _molecule_index(self):
        if self._molecules is None:
            from atomsmm_b200 import engine
            self._molecules = engine._molecules(self.system)[0]
        return self._molecules

    def _barostat_move(self, b):
        force = b['force']
        b['steps'] += 1
        if b['steps'] < force.getFrequency():
            return
        b['steps'] = 0
        seed = force.getRandomNumberSeed() & ((1 << 64) - 1)
        pressure = force.getDefaultPressure().value_in_unit(force.getDefaultPressure().unit)*6.02214179e23*1e-25
        kT = 8.314472471220217e-3*force.getDefaultTemperature().value_in_md_units()
        volume = float(np.prod(self.box))
        if b['scale'] == 0.0:
            b['scale'] = 0.01*volume
        e0 = refmath.evaluate_system(self.system, self.x, self.box, None, self.parameters).energy
        dv = b['scale']*2.0*(self.barostat_uniform(seed, b['counter']) - 0.5)
        b['counter'] += 1
        new_volume = volume + dv
        scale = (new_volume/volume)**(1.0/3.0)
        mol = self._molecule_index()
        nmol = int(mol.max()) + 1
        counts = np.bincount(mol, minlength=nmol)[:, None]
        centres = np.stack([np.bincount(mol, self.x[:, k], nmol) for k in range(3)], axis=1)/counts
        trial = self.x + (centres*(scale - 1.0))[mol]
        new_box = self.box*scale
        e1 = refmath.evaluate_system(self.system, trial, new_box, None, self.parameters).energy
        w = e1 - e0 + pressure*dv - nmol*kT*np.log(new_volume/volume)
        accept = True
        if w > 0:
            accept = self.barostat_uniform(seed, b['counter']) <= np.exp(-w/kT)
        b['counter'] += 1
        if accept:
            self.x, self.box = trial, new_box
            self._version += 1
            self._force_cache = {}
            b['accepted'] += 1
        b['attempts'] += 1
        b['log'].append((bool(accept), float(np.prod(self.box)), float(w)))
        if b['attempts'] >= 10:
            if b['accepted'] < 0.25*b['attempts']:
                b['scale'] /= 1.1
                b['attempts'] = b['accepted'] = 0
            elif b['accepted'] > 0.75*b['attempts']:
                b['scale'] = min(b['scale']*1.1, 0.3*float(np.prod(self.box)))
                b['attempts'] = b['accepted'] = 0

    def step(self, count=1):
        massive = self.mass > 0
        for _ in range(count):
            self.x_constrained = self.x.copy()     # OpenMM's oldPos: reference of the next position constraint
            pc = 0
            loops = []
            while pc < len(self.steps):
                kind, variable, expression = self.steps[pc]
                if kind == 1:      # per DOF
                    value = np.broadcast_to(self._evaluate(expression, True), (self.n, 3)).astype(np.float64)
                    if variable == 'x':
                        self.x = np.where(massive, value, self.x)
                        self._version += 1
                    elif variable == 'v':
                        self.v = np.where(massive, value, self.v)
                    else:
                        self.perdof[variable] = value.copy()
                elif kind == 0:    # global
                    value = float(self._evaluate(expression, False))
                    if variable in self.globals:
                        self.globals[variable] = value
                    else:
                        if self.parameters.get(variable) != value:
                            self._force_cache = {}      # forces depend on context parameters
                        self.parameters[variable] = value
                elif kind == 2:    # sum
                    value = np.broadcast_to(self._evaluate(expression, True), (self.n, 3))
                    self.globals[variable] = float(np.sum(value))
                elif kind == 5:    # UpdateContextState
                    for barostat in self.barostats:
                        self._barostat_move(barostat)
                elif kind == 3:
                    if self.system.getNumConstraints() > 0:
                        self.x = shake(self.constraints(), self.mass[:, 0], self.x, self.x_constrained)
                        self.x_constrained = self.x.copy()
                        self._version += 1
                elif kind == 4:
                    if self.system.getNumConstraints() > 0:
                        self.v = rattle(self.constraints(), self.mass[:, 0], self.x, self.v)
                elif kind == 6:
                    if not self._condition(expression):
                        pc = self._end[pc]
                elif kind == 7:
                    if self._condition(expression):
                        loops.append(pc)
                    else:
                        pc = self._end[pc]
                elif kind == 8:
                    start = [s for s, e in self._end.items() if e == pc][0]
                    if self.steps[start][0] == 7:
                        pc = start - 1
                pc += 1

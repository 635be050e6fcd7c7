"""
Force descriptions for near/far-split Lennard-Jones + Coulomb pair interactions.

Same public names, constructor arguments and emitted energy semantics as the reference's
``atomsmm.forces`` (reference: src/atomsmm/forces.py; class map in SURVEY 8a rows a1-a7).
A force here is a *description* (energy string + parameters on the ``mm`` containers); the
engine recognises the closed family each string belongs to (``lowering.py``) and runs the
matching hand-written sm_100a pair kernel.

Potentials (r in nm, energies in kJ/mol):

* LJC(r)   = 4 eps [(sig/r)^12 - (sig/r)^6] + Kc q1 q2 / r, Lorentz-Berthelot mixing.
* S(u)     = 1 + u^3 (15u - 6u^2 - 10),  u = (r - rs0)/(rc0 - rs0) for r > rs0 else 0.
* near/None          S * LJC                                   (forces.py:541-543)
* near/shift         S * (LJC(r) - LJC(rc0))                   (forces.py:544-548)
* near/force-switch  V*(r) - V*(rc0) with dV/dr = S dLJC/dr    (forces.py:549-563)
* damped-smoothed    [LJ + erfc(alpha r) Kc q1 q2 / r] * switch (forces.py:445-466)
* exceptions         4 eps x (x-1) + Kc qq / r, x = (sig/r)^6   (forces.py:405-407)
"""

import math

from . import mm
from . import unit
from .unit import md_value as _md
from .utils import InputError
from .utils import exceptionOffsetParameters
from .utils import particleOffsetParameters

COULOMB_CONSTANT = 138.935456   # kJ nm / (mol e^2), the literal the reference uses
_KC_QUANTITY = COULOMB_CONSTANT*unit.kilojoules_per_mole/unit.nanometer

_PROPERTIES = ('charge', 'sigma', 'epsilon')
_BOND_PROPERTIES = ('chargeprod', 'sigma', 'epsilon')


class _AtomsMM_Force(object):
    """Mixin: a single description that can add itself to a System (forces.py:24-36)."""

    def addTo(self, system):
        system.addForce(self)


class _AtomsMM_CompoundForce(object):
    """Several descriptions handled as one force (forces.py:39-131).

    Any method not defined here is broadcast to every member that has it and the compound
    object is returned, so calls can be chained exactly as with the reference class.
    """

    def __init__(self, forces):
        self.forces = list(forces) if isinstance(forces, (list, tuple)) else [forces]
        self.setForceGroup(0)

    def __iter__(self):
        return iter(self.forces)

    def __getitem__(self, index):
        return self.forces[index]

    def __len__(self):
        return len(self.forces)

    def __getattr__(self, method):
        if method.startswith('__'):
            raise AttributeError(method)

        def broadcast(*args, **kwargs):
            for member in self.forces:
                target = getattr(member, method, None)
                if target is not None:
                    target(*args, **kwargs)
            return self
        return broadcast

    def getForceGroup(self):
        return self.forces[0].getForceGroup()

    def addTo(self, system):
        for member in self.forces:
            system.addForce(member)
        return self

    def enableExceptions(self):
        extra = NonbondedExceptionsForce()
        extra.setForceGroup(self.getForceGroup())
        self.forces.append(extra)
        return self


class _AtomsMM_NonbondedForce(mm.NonbondedForce, _AtomsMM_Force):
    """Built-in LJ + Coulomb force whose exceptions are all silenced (forces.py:134-190)."""

    def __init__(self, cutoff_distance, switch_distance=None):
        super().__init__()
        self.setCutoffDistance(cutoff_distance)
        self.setUseSwitchingFunction(switch_distance is not None)
        if switch_distance is not None:
            self.setSwitchingDistance(switch_distance)

    def importFrom(self, force):
        for k in range(force.getNumParticles()):
            self.addParticle(*force.getParticleParameters(k))
        for k in range(force.getNumExceptions()):
            i, j, _, sigma, _ = force.getExceptionParameters(k)
            self.addException(i, j, 0.0, sigma, 0.0)
        self.setNonbondedMethod(force.getNonbondedMethod())
        self.setEwaldErrorTolerance(force.getEwaldErrorTolerance())
        self.setPMEParameters(*force.getPMEParameters())
        self.setUseDispersionCorrection(force.getUseDispersionCorrection())
        self.setReactionFieldDielectric(force.getReactionFieldDielectric())
        return self


class _AtomsMM_CustomNonbondedForce(mm.CustomNonbondedForce, _AtomsMM_Force):
    """Pair potential given as a string of r, chargeprod, sigma, epsilon (forces.py:193-323).

    Cutoff, switching and long-range-correction settings left as ``None`` are taken from the
    NonbondedForce handed to :meth:`importFrom`.
    """

    def __init__(self, energy, cutoff_distance=None, use_switching_function=None,
                 switch_distance=None, use_dispersion_correction=None, **global_parameters):
        super().__init__(energy)
        self._deferred = dict(cutoff=cutoff_distance is None, use_switch=use_switching_function is None,
                              switch=switch_distance is None, lrc=use_dispersion_correction is None)
        for name, value in global_parameters.items():
            self.addGlobalParameter(name, value)
        for name in _PROPERTIES:
            self.addPerParticleParameter(name)
        if cutoff_distance is not None:
            self.setCutoffDistance(cutoff_distance)
        if use_switching_function is not None:
            self.setUseSwitchingFunction(use_switching_function)
        if switch_distance is not None:
            self.setSwitchingDistance(switch_distance)
        if use_dispersion_correction is not None:
            self.setUseLongRangeCorrection(use_dispersion_correction)

    def __repr__(self):
        return '\n'.join(term.strip(' \t') for term in self.getEnergyFunction().split(';'))

    def mixingRules(self, offset_parameters):
        """``;chargeprod = ..;sigma = ..;epsilon = ..`` with optional lambda-scaled offsets."""
        side = {}
        for tag in ('1', '2'):
            side[tag] = {p: p + tag for p in _PROPERTIES}
            for parameter in offset_parameters:
                for p in _PROPERTIES:
                    side[tag][p] += '+{0}*{1}Scale_{0}{2}'.format(parameter, p, tag)
        rules = ';chargeprod = ({})*({})'.format(side['1']['charge'], side['2']['charge'])
        rules += ';sigma = 0.5*({}+{})'.format(side['1']['sigma'], side['2']['sigma'])
        rules += ';epsilon = sqrt(({})*({}))'.format(side['1']['epsilon'], side['2']['epsilon'])
        return rules

    def importFrom(self, nonbonded):
        builtin, custom = mm.NonbondedForce, mm.CustomNonbondedForce
        periodic = (builtin.CutoffPeriodic, builtin.Ewald, builtin.PME, builtin.LJPME)
        method = nonbonded.getNonbondedMethod()
        if method in periodic:
            self.setNonbondedMethod(custom.CutoffPeriodic)
        elif method == builtin.CutoffNonPeriodic:
            self.setNonbondedMethod(custom.CutoffNonPeriodic)
        else:
            self.setNonbondedMethod(custom.NoCutoff)
        if self._deferred['cutoff']:
            self.setCutoffDistance(nonbonded.getCutoffDistance())
        if self._deferred['use_switch']:
            self.setUseSwitchingFunction(nonbonded.getUseSwitchingFunction())
        if self._deferred['switch']:
            self.setSwitchingDistance(nonbonded.getSwitchingDistance())
        if self._deferred['lrc']:
            self.setUseLongRangeCorrection(nonbonded.getUseDispersionCorrection())
        offsets = particleOffsetParameters(nonbonded)
        self.setEnergyFunction(self.getEnergyFunction() + self.mixingRules(offsets))
        for parameter, value in offsets.items():
            self.addGlobalParameter(parameter, value)
            for p in _PROPERTIES:
                self.addPerParticleParameter('{}Scale_{}'.format(p, parameter))
        padding = [0.0]*(3*len(offsets))
        for k in range(nonbonded.getNumParticles()):
            self.addParticle(list(nonbonded.getParticleParameters(k)) + padding)
        column = {name: 3*(n + 1) for n, name in enumerate(offsets)}
        for k in range(nonbonded.getNumParticleParameterOffsets()):
            parameter, particle, *scales = nonbonded.getParticleParameterOffset(k)
            values = list(self.getParticleParameters(particle))
            values[column[parameter]:column[parameter] + 3] = scales
            self.setParticleParameters(particle, values)
        for k in range(nonbonded.getNumExceptions()):
            i, j = nonbonded.getExceptionParameters(k)[:2]
            self.addExclusion(i, j)
        return self

    def getGlobalParameters(self):
        return {self.getGlobalParameterName(k): self.getGlobalParameterDefaultValue(k)
                for k in range(self.getNumGlobalParameters())}


class _AtomsMM_CustomBondForce(mm.CustomBondForce, _AtomsMM_Force):
    """Pair potential evaluated on an explicit pair list (forces.py:326-397)."""

    def __init__(self, energy, **globalParams):
        super().__init__(energy)
        for name, value in globalParams.items():
            self.addGlobalParameter(name, value)

    def offsetRules(self, offset_parameters):
        rules = ''
        for p in _BOND_PROPERTIES:
            terms = p + '0'
            for parameter in offset_parameters:
                terms += '+{0}*{1}Scale_{0}'.format(parameter, p)
            rules += ';{} = {}'.format(p, terms)
        return rules

    def importFrom(self, nonbonded, extract=False):
        self.setUsesPeriodicBoundaryConditions(nonbonded.usesPeriodicBoundaryConditions())
        offsets = exceptionOffsetParameters(nonbonded)
        if offsets:
            self.setEnergyFunction(self.getEnergyFunction() + self.offsetRules(offsets))
            for p in _BOND_PROPERTIES:
                self.addPerBondParameter(p + '0')
            for parameter, value in offsets.items():
                self.addGlobalParameter(parameter, value)
                for p in _BOND_PROPERTIES:
                    self.addPerBondParameter('{}Scale_{}'.format(p, parameter))
        else:
            for p in _BOND_PROPERTIES:
                self.addPerBondParameter(p)
        padding = [0.0]*(3*len(offsets))
        for k in range(nonbonded.getNumExceptions()):
            i, j, chargeprod, sigma, epsilon = nonbonded.getExceptionParameters(k)
            self.addBond(i, j, [chargeprod, sigma, epsilon] + padding)
            if extract:
                nonbonded.setExceptionParameters(k, i, j, 0.0, 1.0, 0.0)
        column = {name: 3*(n + 1) for n, name in enumerate(offsets)}
        for k in range(nonbonded.getNumExceptionParameterOffsets()):
            parameter, bond, *scales = nonbonded.getExceptionParameterOffset(k)
            i, j, values = self.getBondParameters(bond)
            values = list(values)
            values[column[parameter]:column[parameter] + 3] = scales
            self.setBondParameters(bond, i, j, values)
        return self


_LJ = '4*epsilon*((sigma/r)^12-(sigma/r)^6)'
_EXCEPTION_LJC = '4*epsilon*x*(x-1) + Kc*chargeprod/r; x=(sigma/r)^6'


class NonbondedExceptionsForce(_AtomsMM_CustomBondForce):
    """All exceptions of a NonbondedForce as an explicit-pair LJC force (forces.py:400-407)."""

    def __init__(self):
        super().__init__(_EXCEPTION_LJC, Kc=_KC_QUANTITY)


class DampedSmoothedForce(_AtomsMM_CustomNonbondedForce):
    """LJ + erfc-damped Coulomb, smoothly switched off in [rswitch, rcut] (forces.py:410-466).

    ``degree`` 1 uses the built-in 5th-order switch in r; ``degree`` d >= 2 uses the same
    polynomial in u = (r^d - rswitch^d)/(rcut^d - rswitch^d).  No long-range correction.
    """

    def __init__(self, alpha, cutoff_distance, switch_distance, degree=1):
        rs, rc = _md(switch_distance), _md(cutoff_distance)
        if rs < 0.0 or rs >= rc:
            raise InputError('Switching distance must satisfy 0 <= r_switch < r_cutoff')
        core = '4*epsilon*((sigma/r)^12 - (sigma/r)^6) + erfc(alpha*r)*Kc*chargeprod/r'
        linear = degree == 1
        if linear:
            energy = core
        else:
            energy = ';'.join(['S*({})'.format(core),
                               'S = 1 + step(r - rswitch)*u^3*(15*u - 6*u^2 - 10)',
                               'u = (r^d - rswitch^d)/(rcut^d - rswitch^d); d={}'.format(degree)])
        super().__init__(energy=energy, cutoff_distance=cutoff_distance, use_switching_function=linear,
                         switch_distance=switch_distance if linear else None,
                         use_dispersion_correction=False, Kc=_KC_QUANTITY, alpha=alpha,
                         rswitch=switch_distance, rcut=cutoff_distance)


# ---------------------------------------------------------------------------------------------
# Near-force expression builders
# ---------------------------------------------------------------------------------------------

_SWITCH = 'S = 1 + step(r - rs0)*u^3*(15*u - 6*u^2 - 10)'
_U = 'u=(r-rs0)/(rc0-rs0)'

# f_n(u) solve f_n - (u+b)/n f_n' = S(u), f_n(0) = 1, so that d/dr [f_n(u)/r^n] = S(u) d/dr r^-n.
_FS_FACTORS = (
    ('f12', '(6*b^2-21*b+28)*(b^3*(R^12-1)-12*b^2*u-66*b*u^2-220*u^3)/462+45*(7-2*b)*u^4/14-72*u^5/7'),
    ('f6', '(6*b^2-3*b+1)*(b^3*(R^6-1)-6*b^2*u-15*b*u^2-20*u^3)+45*(1-2*b)*u^4-36*u^5'),
    ('f1', '5*(b+1)^2*(6*b^3*R*log(R)-6*b^2*u-3*b*u^2+u^3)+u^4*(3*u-5*b-10)/2'),
)


def force_switch_constants(rs, rc, b=None):
    """b = rs/(rc-rs) and the values f12(1), f6(1), f1(1) at the cutoff (forces.py:489-493)."""
    if b is None:
        b = rs/(rc - rs)
    f12c = (1 + b)**3*(b**6 + 3*b**5 + (30/7)*b**4 + (25/7)*b**3 + (25/14)*b**2 + (1/2)*b + 2/33)/b**9
    f6c = (1 + b)**3/b**3
    f1c = (30*(1 + b))*(b**2*(1 + b)**2*math.log(1/b + 1) - b**3 - (3/2)*b**2 - (1/3)*b + 1/12)
    return b, f12c, f6c, f1c


def _b_ratio(cutoff_distance, switch_distance):
    """rs/(rc-rs) evaluated in the units the caller used (the reference does the same, so e.g.
    9.5 A / 0.5 A gives exactly 19 whereas 0.95 nm / 0.05 nm does not)."""
    return float(switch_distance/(cutoff_distance - switch_distance))


def _near_terms(rs, rc, adjustment, coulomb=True, tail=(), b=None):
    """Main term + auxiliary definitions of a near potential, as a list of strings."""
    if adjustment is None:
        body = _LJ + (' + Kc*chargeprod/r' if coulomb else '')
        terms = ['S*({})'.format(body), _SWITCH]
    elif adjustment == 'shift':
        body = '4*epsilon*((sigma/r)^12-(sigma/r)^6-((sigma/rc0)^12-(sigma/rc0)^6))'
        if coulomb:
            body = '{}+{}'.format(body, 'Kc*chargeprod*(1/r-1/rc0)')
        terms = ['S*({})'.format(body), _SWITCH]
    elif adjustment == 'force-switch':
        at_r = '4*epsilon*(f12*(sigma/r)^12-f6*(sigma/r)^6)'
        at_rc = '4*epsilon*(f12c*(sigma/rc0)^12-f6c*(sigma/rc0)^6)'
        if coulomb:
            at_r += ' + Kc*chargeprod*f1/r'
            at_rc += ' + Kc*chargeprod*f1c/rc0'
        terms = ['{}-({})'.format(at_r, at_rc)]
        b, f12c, f6c, f1c = force_switch_constants(rs, rc, b)
        for name, poly in _FS_FACTORS[:3 if coulomb else 2]:
            terms.append('{}=1+step(r-rs0)*({})'.format(name, poly))
        terms += ['R=u/b+1', 'b={}'.format(b), 'f12c={}'.format(f12c), 'f6c={}'.format(f6c)]
        if coulomb:
            terms.append('f1c={}'.format(f1c))
    else:
        raise InputError('unknown adjustment option')
    terms.append(_U)
    terms.extend(tail)
    return terms


def nearForceExpressions(cutoff_distance, switch_distance, adjustment):
    """Near LJC potential with literal rs0/rc0/Kc definitions (forces.py:469-500).

    With ``adjustment=None`` the reference prefixes the main term with ``energy=``; the prefix
    carries no meaning for the evaluator and is kept for string parity.
    """
    rs, rc = _md(switch_distance), _md(cutoff_distance)
    terms = _near_terms(rs, rc, adjustment, True, ['rs0={}'.format(rs), 'rc0={}'.format(rc), 'Kc=138.935456'],
                        _b_ratio(cutoff_distance, switch_distance))
    if adjustment is None:
        terms[0] = 'energy=' + terms[0]
    return terms


def nearLJForceExpressions(cutoff_distance, switch_distance, adjustment):
    """Same without the Coulomb part (forces.py:503-530)."""
    rs, rc = _md(switch_distance), _md(cutoff_distance)
    terms = _near_terms(rs, rc, adjustment, False, ['rs0={}'.format(rs), 'rc0={}'.format(rc)],
                        _b_ratio(cutoff_distance, switch_distance))
    if adjustment is None:
        terms[0] = 'energy=' + terms[0]
    return terms


class NearForce(object):
    def _globalParams(self, cutoff_distance, switch_distance):
        return {'Kc': _KC_QUANTITY, 'rc0': cutoff_distance, 'rs0': switch_distance}

    def _expressions(self, cutoff_distance, switch_distance, adjustment):
        return _near_terms(_md(switch_distance), _md(cutoff_distance), adjustment,
                           b=_b_ratio(cutoff_distance, switch_distance))


class NearNonbondedForce(_AtomsMM_CustomNonbondedForce, NearForce):
    """Short-range, smoothly truncated LJC pair force for the fast RESPA levels
    (forces.py:570-670).

    Parameters
    ----------
        cutoff_distance, switch_distance : unit.Quantity
            rc0 and rs0 above.
        adjustment : None | 'shift' | 'force-switch'
        subtract : bool
            Emit minus the potential.
        actual_cutoff : unit.Quantity, optional
            Cutoff the neighbour search really uses; the potential is then gated by
            ``step(rc0-r)``.
    """

    def __init__(self, cutoff_distance, switch_distance, adjustment=None, subtract=False,
                 actual_cutoff=None):
        terms = self._expressions(cutoff_distance, switch_distance, adjustment)
        if actual_cutoff is not None:
            terms[0] = 'step(rc0-r)*({})'.format(terms[0])
        if subtract:
            terms[0] = '-({})'.format(terms[0])
        super().__init__(energy='; '.join(terms),
                         cutoff_distance=cutoff_distance if actual_cutoff is None else actual_cutoff,
                         use_switching_function=False, use_dispersion_correction=False,
                         **self._globalParams(cutoff_distance, switch_distance))


class NearExceptionForce(_AtomsMM_CustomBondForce, NearForce):
    """The near potential on an explicit pair list, gated at rc0 (forces.py:673-680)."""

    def __init__(self, cutoff_distance, switch_distance, adjustment=None, subtract=False):
        terms = self._expressions(cutoff_distance, switch_distance, adjustment)
        terms[0] = '{}step(rc0-r)*({})'.format('-' if subtract else '', terms[0])
        super().__init__('; '.join(terms), **self._globalParams(cutoff_distance, switch_distance))


class FarNonbondedForce(_AtomsMM_CompoundForce):
    """Complement of a :class:`NearNonbondedForce`: full LJ + Coulomb within the outer cutoff
    minus the near potential inside rc0 (forces.py:683-724)."""

    def __init__(self, preceding, cutoff_distance, switch_distance=None):
        if not isinstance(preceding, NearNonbondedForce):
            raise InputError('argument \'preceding\' must be of class NearNonbondedForce')
        pieces = preceding.getEnergyFunction().split(';')
        pieces[0] = '-step(rc0-r)*({})'.format(pieces[0])
        discount = _AtomsMM_CustomNonbondedForce(energy=';'.join(pieces), cutoff_distance=cutoff_distance,
                                                 use_switching_function=False, use_dispersion_correction=False,
                                                 **preceding.getGlobalParameters())
        total = _AtomsMM_NonbondedForce(cutoff_distance, switch_distance)
        super().__init__([total, discount])


class SoftcoreLennardJonesForce(_AtomsMM_CustomNonbondedForce):
    """Beutler soft-core LJ, V = 4 lambda eps x (x-1), x = 1/((r/sig)^6 + (1-lambda)/2)
    (forces.py:727-758)."""

    def __init__(self, cutoff_distance=None, use_switching_function=None, switch_distance=None,
                 use_dispersion_correction=None, parameter='lambda'):
        energy = '4*{0}*epsilon*x*(x-1);x = 1/((r/sigma)^6 + 0.5*(1-{0}))'.format(parameter)
        super().__init__(energy=energy, cutoff_distance=cutoff_distance,
                         use_switching_function=use_switching_function, switch_distance=switch_distance,
                         use_dispersion_correction=use_dispersion_correction, **{parameter: 1.0})


class SoftcoreForce(_AtomsMM_CustomNonbondedForce):
    """Soft-core LJ + linearly scaled Coulomb with lambda_vdw / lambda_coul (forces.py:761-793)."""

    def __init__(self, cutoff_distance, switch_distance=None):
        energy = '4*lambda_vdw*epsilon*(1-x)/x^2 + Kc*lambda_coul*chargeprod/r;'
        energy += 'x = (r/sigma)^6 + 0.5*(1-lambda_vdw)'
        super().__init__(energy=energy, cutoff_distance=cutoff_distance, use_switching_function=True,
                         switch_distance=switch_distance, Kc=_KC_QUANTITY, lambda_vdw=1.0, lambda_coul=1.0)

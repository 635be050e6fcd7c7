"""
Propagators: composable exponential-operator building blocks that *emit step programs*.

Same class names, constructor arguments and composition algebra as the reference's
``atomsmm.propagators`` (reference: src/atomsmm/propagators.py; rows a11-a17, a20 of
SURVEY 8a).  A propagator appends ``(kind, variable, expression)`` steps, scaled by a
``fraction`` of the time step, to an ``integrators._AtomsMM_Integrator``; the engine lowers
that program to fused CUDA kernels (``lowering.py`` / csrc/integrate.cu).  The emitted
programs are checked against programs captured from the reference's own Python layer
(tests/golden/ref_programs.json).

Deliberate, documented deviations:

* ``VelocityRescalingPropagator`` (Bussi): the reference draws ``gaussian`` inside a per-DOF
  expression, which under CustomIntegrator semantics gives every degree of freedom its own
  "global" scaling factor, and reads an undefined variable ``random``
  (propagators.py:1210,1219-1227).  The default here implements the published algorithm the
  reference's docstring states: one normal + one gamma deviate per application and a single
  global factor.  ``per_dof_noise=True`` reproduces the reference's literal program.
* ``NoseHooverPropagator``: the sub-loop guard ``nloops > 2`` (propagators.py:1264) means
  ``nloops == 2`` integrates half the interval; preserved for program parity.
"""

import math
import re

from . import mm
from . import unit
from .unit import md_value as _md
from .utils import InputError
from .utils import kB


class Propagator(object):
    """Base class (propagators.py:24-75): owns the global / per-DOF variables it needs."""

    def __init__(self):
        self.globalVariables = dict()
        self.perDofVariables = dict()

    def addVariables(self, integrator):
        for name, value in self.globalVariables.items():
            integrator.addGlobalVariable(name, value)
        for name, value in self.perDofVariables.items():
            integrator.addPerDofVariable(name, value)

    def absorbVariables(self, propagator):
        for mine, theirs, what in ((self.globalVariables, propagator.globalVariables, 'Global'),
                                   (self.perDofVariables, propagator.perDofVariables, 'Per-dof')):
            for key, value in theirs.items():
                if key in mine and value != mine[key]:
                    raise InputError('%s variable inconsistency in merged propagators' % what)
            mine.update(theirs)

    def addSteps(self, integrator, fraction=1.0, force='f'):
        pass

    def integrator(self, stepSize):
        """An integrator that applies this propagator once per step of size ``stepSize``."""
        from . import integrators
        new = integrators._AtomsMM_Integrator(stepSize)
        self.addVariables(new)
        self.addSteps(new)
        return new


class ChainedPropagator(Propagator):
    """Apply a list of propagators one after another (propagators.py:78-114)."""

    def __init__(self, propagators):
        super().__init__()
        self.propagators = propagators
        for member in propagators:
            self.absorbVariables(member)

    def addSteps(self, integrator, fraction=1.0, force='f'):
        for member in self.propagators:
            member.addSteps(integrator, fraction, force)


class SplitPropagator(Propagator):
    """``A(dt) = [A(dt/n)]^n`` as a counted while-loop (propagators.py:117-149)."""

    def __init__(self, A, n):
        super().__init__()
        self.A, self.n = A, n
        self.absorbVariables(A)
        self.globalVariables['nSplit'] = 0

    def addSteps(self, integrator, fraction=1.0, force='f'):
        if self.n == 1:
            self.A.addSteps(integrator, fraction, force)
            return
        integrator.addComputeGlobal('nSplit', '0')
        integrator.beginWhileBlock('nSplit < {}'.format(self.n))
        self.A.addSteps(integrator, fraction/self.n)
        integrator.addComputeGlobal('nSplit', 'nSplit + 1')
        integrator.endBlock()


class TrotterSuzukiPropagator(Propagator):
    """Symmetric splitting ``B(dt/2) A(dt) B(dt/2)`` (propagators.py:152-187)."""

    def __init__(self, A, B):
        super().__init__()
        self.A, self.B = A, B
        self.absorbVariables(A)
        self.absorbVariables(B)

    def addSteps(self, integrator, fraction=1.0, force='f'):
        self.B.addSteps(integrator, 0.5*fraction, force)
        self.A.addSteps(integrator, fraction, force)
        self.B.addSteps(integrator, 0.5*fraction, force)


_SY_WEIGHTS = {
    1: [],
    3: [1.3512071919596578],
    7: [0.784513610477560, 0.235573213359357, -1.17767998417887],
    15: [0.9148442462, 0.2536933366, -1.4448522369, -0.1582406354, 1.9381391376, -1.960610233, 0.1027998494],
}


class SuzukiYoshidaPropagator(Propagator):
    """Higher-order symmetric factorisation with 1, 3, 7 or 15 weights (propagators.py:190-226)."""

    def __init__(self, A, nsy=3):
        super().__init__()
        if nsy not in _SY_WEIGHTS:
            raise InputError('SuzukiYoshidaPropagator accepts nsy = 1, 3, 7, or 15 only')
        self.A, self.nsy = A, nsy
        self.absorbVariables(A)

    def addSteps(self, integrator, fraction=1.0, force='f'):
        half = _SY_WEIGHTS[self.nsy]
        for w in half + [1 - 2*sum(half)] + half[::-1]:
            self.A.addSteps(integrator, fraction*w)


class TranslationPropagator(Propagator):
    """Drift ``x += (fraction*dt) v``, optionally SHAKE-constrained (propagators.py:229-252)."""

    def __init__(self, constrained=True):
        super().__init__()
        self.constrained = constrained
        if constrained:
            self.perDofVariables['x0'] = 0

    def addSteps(self, integrator, fraction=1.0, force='f'):
        if self.constrained:
            integrator.addComputePerDof('x0', 'x')
        integrator.addComputePerDof('x', 'x + ({}*dt)*v'.format(fraction))
        if self.constrained:
            integrator.addConstrainPositions()
            integrator.addComputePerDof('v', '(x - x0)/({}*dt)'.format(fraction))


class VelocityBoostPropagator(Propagator):
    """Kick ``v += (fraction*dt) F/m``, optionally RATTLE-constrained (propagators.py:255-273)."""

    def __init__(self, constrained=True):
        super().__init__()
        self.constrained = constrained

    def addSteps(self, integrator, fraction=1.0, force='f'):
        integrator.addComputePerDof('v', 'v + ({}*dt)*{}/m'.format(fraction, force))
        if self.constrained:
            integrator.addConstrainVelocities()


class OrnsteinUhlenbeckPropagator(Propagator):
    """Exact solution of ``dV = F/M dt - gamma V dt + sqrt(2 gamma kT/M) dW`` over
    ``fraction*dt`` (propagators.py:685-741)."""

    def __init__(self, temperature, frictionConstant, velocity='v', mass='m', force=None,
                 overall=False, **globals):
        super().__init__()
        self.globalVariables['kT'] = kB*temperature
        self.globalVariables['friction'] = frictionConstant
        self.velocity, self.mass, self.force, self.overall = velocity, mass, force, overall
        self.globalVariables.update(globals)
        if velocity != 'v':
            (self.globalVariables if overall else self.perDofVariables)[velocity] = 0

    def addSteps(self, integrator, fraction=1.0, force='f'):
        expression = 'z*{} + sqrt(kT*(1 - z*z)/mass)*gaussian'.format(self.velocity)
        if self.force is not None:
            expression += ' + force*(1 - z)/(mass*friction); force = {}'.format(self.force)
        expression += '; mass = {}'.format(self.mass)
        expression += '; z = exp(-({}*dt)*friction)'.format(fraction)
        if self.overall:
            integrator.addComputeGlobal(self.velocity, expression)
        else:
            integrator.addComputePerDof(self.velocity, expression)


class GenericBoostPropagator(Propagator):
    """``dV/dt = F/M`` for named variables (propagators.py:744-790)."""

    def __init__(self, velocity='v', mass='m', force='f', perDof=True, **globals):
        super().__init__()
        self.velocity, self.mass, self.force, self.perDof = velocity, mass, force, perDof
        self.globalVariables.update(globals)
        if velocity != 'v':
            (self.perDofVariables if perDof else self.globalVariables)[velocity] = 0

    def addSteps(self, integrator, fraction=1.0, force='f'):
        expression = '{} + ({}*dt)*F/M; F = {}; M = {}'.format(self.velocity, fraction, self.force, self.mass)
        add = integrator.addComputePerDof if self.perDof else integrator.addComputeGlobal
        add(self.velocity, expression)


class GenericScalingPropagator(Propagator):
    """``dV/dt = -damping*V`` for named variables (propagators.py:793-827)."""

    def __init__(self, velocity, damping, perDof=True, **globals):
        super().__init__()
        self.velocity, self.damping, self.perDof = velocity, damping, perDof
        self.globalVariables.update(globals)
        if perDof and velocity != 'v':
            self.perDofVariables[velocity] = 0
        elif not perDof:
            self.globalVariables[velocity] = 0

    def addSteps(self, integrator, fraction=1.0, force='f'):
        expression = '{}*exp(-({}*dt)*{})'.format(self.velocity, fraction, self.damping)
        add = integrator.addComputePerDof if self.perDof else integrator.addComputeGlobal
        add(self.velocity, expression)


class RespaPropagator(Propagator):
    """Multiple-time-scale rRESPA propagator over force groups 0..N-1 (propagators.py:830-973).

    Level k (force group k, ``loops[k]`` iterations per iteration of level k+1) is

        [ shell_k(h/2) boost_k(h/2) LEVEL_{k-1}(h) boost_k(h/2) shell_k(h/2) ]^loops[k]

    and the innermost level is ``move(h)`` or ``move(h/2) core(h) move(h/2)``.  Group 0 and 1
    are kicked with ``f0``/``f1``; group k >= 2 with the *difference* ``f{k} - f{k-1}`` so that
    group k may hold the full interaction of which group k-1 is the short-range part.

    Keyword arguments: ``has_memory`` (default False), ``use_respa_switch``, ``blitz``.
    """

    def __init__(self, loops, move=None, boost=None, core=None, shell=None, **kwargs):
        super().__init__()
        self.loops = loops
        self.N = len(loops)
        self.move = TranslationPropagator(constrained=False) if move is None else move
        self.boost = VelocityBoostPropagator(constrained=False) if boost is None else boost
        self.core = core
        self.shell = dict() if shell is None else shell
        if not set(self.shell).issubset(range(self.N)):
            raise InputError('invalid key(s) in RespaPropagator \'shell\' argument')
        for member in [self.move, self.boost, self.core] + list(self.shell.values()):
            if member is not None:
                self.absorbVariables(member)
        for level, n in enumerate(loops):
            if n > 1:
                self.globalVariables['n{}RESPA'.format(level)] = 0
        self.expr = ['f{}'.format(k) if k < 2 else 'f{}-f{}'.format(k, k - 1) for k in range(self.N)]
        self.force = list(self.expr)
        self._has_memory = kwargs.pop('has_memory', False)
        if self._has_memory:
            for k in range(1, self.N):
                self.perDofVariables['fm{}'.format(k)] = 0.0
                self.force[0] += '+fm{}'.format(k)
                self.force[k] += '-fm{}'.format(k)
        self.force = ['({})'.format(f) for f in self.force]
        self._use_respa_switch = kwargs.pop('use_respa_switch', False)
        self._blitz = kwargs.pop('blitz', False)

    def addSteps(self, integrator, fraction=1.0, force='f'):
        if self._use_respa_switch:
            integrator.addComputeGlobal('respa_switch', '1')
        self._addSubsteps(integrator, self.N - 1, fraction)
        if self._use_respa_switch:
            integrator.addComputeGlobal('respa_switch', '0')

    def _internalSplitting(self, integrator, timescale, fraction, shell):
        remembered = self._has_memory and timescale > 0
        if self._blitz:
            if remembered:
                integrator.addComputePerDof('F{}'.format(timescale), 'f{}'.format(timescale))
            else:
                self.boost.addSteps(integrator, fraction, self.force[timescale])
            self._addSubsteps(integrator, timescale - 1, fraction)
            return
        if shell is not None:
            shell.addSteps(integrator, 0.5*fraction, self.force[timescale])
        if remembered:
            integrator.addComputePerDof('fm{}'.format(timescale), self.expr[timescale])
        else:
            self.boost.addSteps(integrator, 0.5*fraction, self.force[timescale])
        self._addSubsteps(integrator, timescale - 1, fraction)
        self.boost.addSteps(integrator, 0.5*fraction, self.force[timescale])
        if shell is not None:
            shell.addSteps(integrator, 0.5*fraction, self.force[timescale])

    def _addSubsteps(self, integrator, timescale, fraction):
        if timescale < 0:
            if self.core is None:
                self.move.addSteps(integrator, fraction)
            else:
                self.move.addSteps(integrator, 0.5*fraction)
                self.core.addSteps(integrator, fraction)
                self.move.addSteps(integrator, 0.5*fraction)
            return
        n = self.loops[timescale]
        counter = 'n{}RESPA'.format(timescale)
        if n > 1:
            integrator.addComputeGlobal(counter, '0')
            integrator.beginWhileBlock('{} < {}'.format(counter, n))
        self._internalSplitting(integrator, timescale, fraction/n, self.shell.get(timescale, None))
        if n > 1:
            integrator.addComputeGlobal(counter, '{} + 1'.format(counter))
            integrator.endBlock()


class MultipleTimeScalePropagator(RespaPropagator):
    """RESPA with a bath placed by ``scheme`` in {middle, blitz, xi-respa, xo-respa, side} and
    factorised by ``nres`` / ``nsy`` (propagators.py:976-1042)."""

    def __init__(self, loops, move=None, boost=None, bath=None, **kwargs):
        scheme = kwargs.pop('scheme', 'middle')
        location = kwargs.pop('location', 0)
        nres = kwargs.pop('nres', 1)
        nsy = kwargs.pop('nsy', 1)
        if nres > 1:
            bath = SplitPropagator(bath, nres)
        if nsy > 1:
            bath = SuzukiYoshidaPropagator(bath, nsy)
        if scheme in ('middle', 'blitz'):
            if scheme == 'blitz':
                kwargs['blitz'] = True
            super().__init__(loops, move=move, boost=boost, core=bath, **kwargs)
        elif scheme in ('xi-respa', 'xo-respa', 'side'):
            level = {'side': location, 'xi-respa': 0, 'xo-respa': len(loops) - 1}[scheme]
            super().__init__(loops, move=move, boost=boost, shell={level: bath}, **kwargs)
        else:
            raise InputError('wrong value of scheme parameter')


class VelocityVerletPropagator(Propagator):
    """Velocity Verlet with SHAKE/RATTLE calls (propagators.py:1108-1133)."""

    def __init__(self):
        super().__init__()
        self.perDofVariables['x0'] = 0

    def addSteps(self, integrator, fraction=1.0, force='f'):
        Dt = '; Dt=%s*dt' % fraction
        integrator.addComputePerDof('v', 'v+0.5*Dt*f/m' + Dt)
        integrator.addComputePerDof('x0', 'x')
        integrator.addComputePerDof('x', 'x+Dt*v' + Dt)
        integrator.addConstrainPositions()
        integrator.addComputePerDof('v', '(x-x0)/Dt+0.5*Dt*f/m' + Dt)
        integrator.addConstrainVelocities()


class UnconstrainedVelocityVerletPropagator(Propagator):
    """Plain kick-drift-kick (propagators.py:1136-1153)."""

    def addSteps(self, integrator, fraction=1.0, force='f'):
        integrator.addComputePerDof('v', 'v+0.5*{}*dt*f/m'.format(fraction))
        integrator.addComputePerDof('x', 'x+{}*dt*v'.format(fraction))
        integrator.addComputePerDof('v', 'v+0.5*{}*dt*f/m'.format(fraction))


class VelocityRescalingPropagator(Propagator):
    """Stochastic velocity rescaling of Bussi, Donadio and Parrinello (propagators.py:1156-1227).

    Kinetic energy K relaxes towards N_f kT/2 with time constant ``timeScale`` through

        alpha^2 = A + (kT/mvv) (1-A) (R1^2 + sum_{i=2}^{N_f} R_i^2) + 2 R1 sqrt((kT/mvv) A (1-A)),

    A = exp(-fraction*dt/tau); the chi-square sum is drawn as 2*Gamma((N_f-2+N_f%2)/2)
    (+ one extra normal squared if N_f is odd) with the Marsaglia-Tsang rejection method
    executed inside the step program.
    """

    def __init__(self, temperature, degreesOfFreedom, timeScale, per_dof_noise=False):
        super().__init__()
        self.tau = _md(timeScale)
        self.dof = degreesOfFreedom
        self.kT = _md(unit.BOLTZMANN_CONSTANT_kB*unit.AVOGADRO_CONSTANT_NA*temperature)
        self.per_dof_noise = per_dof_noise
        for name in ('V', 'X', 'U', 'ready'):
            self.globalVariables[name] = 0
        if not per_dof_noise:
            self.globalVariables['R1'] = 0
            self.globalVariables['vscaling'] = 0

    def addSteps(self, integrator, fraction=1.0, force='f'):
        shape = (self.dof - 2 + self.dof % 2)/2
        d = shape - 1/3
        c = 1/math.sqrt(9*d)
        integrator.addComputeGlobal('ready', '0')
        integrator.beginWhileBlock('ready < 0.5')
        integrator.addComputeGlobal('X', 'gaussian')
        integrator.addComputeGlobal('V', '1+%s*X' % c)
        integrator.beginWhileBlock('V <= 0.0')
        integrator.addComputeGlobal('X', 'gaussian')
        integrator.addComputeGlobal('V', '1+%s*X' % c)
        integrator.endBlock()
        integrator.addComputeGlobal('V', 'V^3')
        integrator.addComputeGlobal('U', 'random' if self.per_dof_noise else 'uniform')
        integrator.addComputeGlobal('ready', 'step(1-0.0331*X^4-U)')
        integrator.beginIfBlock('ready < 0.5')
        integrator.addComputeGlobal('ready', 'step(0.5*X^2+%s*(1-V+log(V))-log(U))' % d)
        integrator.endBlock()
        integrator.endBlock()
        odd = self.dof % 2 == 1
        if odd:
            integrator.addComputeGlobal('X', 'gaussian')
        noise = 'gaussian' if self.per_dof_noise else 'R1'
        scaling = 'sqrt(A+C*B*({0}^2+sumRs)+2*sqrt(C*B*A)*{0})'.format(noise)
        scaling += '; C = %s/mvv' % self.kT
        scaling += '; B = 1-A'
        scaling += '; A = exp(-dt*%s)' % (fraction/self.tau)
        # factor 2: chi-square(2a) = 2*Gamma(a); the rejection loop leaves Gamma(a) = d*V
        scaling += '; sumRs = %s*V' % (2*d) + ('+X^2' if odd else '')
        if self.per_dof_noise:
            integrator.addComputePerDof('v', 'vscaling*v; vscaling = ' + scaling)
        else:
            integrator.addComputeGlobal('R1', 'gaussian')
            integrator.addComputeGlobal('vscaling', scaling)
            integrator.addComputePerDof('v', 'vscaling*v')


class NoseHooverPropagator(Propagator):
    """Single global Nose-Hoover thermostat, Q = N_f kT tau^2 (propagators.py:1230-1273):
    half kick of p_eta, velocity scaling exp(-h p_eta/Q), half kick with the scaled mvv."""

    def __init__(self, temperature, degreesOfFreedom, timeScale, nloops=1):
        super().__init__()
        self.nloops = nloops
        self.globalVariables['LkT'] = degreesOfFreedom*kB*temperature
        self.globalVariables['Q'] = degreesOfFreedom*kB*temperature*timeScale**2
        self.globalVariables['vscaling'] = 0
        self.globalVariables['p_eta'] = 0
        self.globalVariables['n_NH'] = 0

    def addSteps(self, integrator, fraction=1.0, force='f'):
        n = self.nloops
        h = fraction/n
        integrator.addComputeGlobal('p_eta', 'p_eta + ({}*dt)*(mvv - LkT)'.format(0.5*h))
        integrator.addComputeGlobal('vscaling', 'exp(-({}*dt)*p_eta/Q)'.format(h))
        if n > 2:
            integrator.addComputeGlobal('n_NH', '1')
            integrator.beginWhileBlock('n_NH < {}'.format(n))
            integrator.addComputeGlobal('p_eta', 'p_eta + ({}*dt)*(vscaling^2*mvv - LkT)'.format(h))
            integrator.addComputeGlobal('vscaling', 'vscaling*exp(-({}*dt)*p_eta/Q)'.format(h))
            integrator.addComputeGlobal('n_NH', 'n_NH + 1')
            integrator.endBlock()
        integrator.addComputeGlobal('p_eta', 'p_eta + ({}*dt)*(vscaling^2*mvv - LkT)'.format(0.5*h))
        integrator.addComputePerDof('v', 'vscaling*v')


class NoseHooverChainPropagator(Propagator):
    """Two-thermostat Nose-Hoover chain, B2 S1 B1 S B1 S1 B2 splitting (propagators.py:1362-1449)."""

    def __init__(self, temperature, degreesOfFreedom, timeScale, frictionConstant=None):
        super().__init__()
        self.temperature, self.degreesOfFreedom, self.timeScale = temperature, degreesOfFreedom, timeScale
        self.frictionConstant = 1/timeScale if frictionConstant is None else frictionConstant
        for name in ('vscaling', 'p_NHC_1', 'p_NHC_2'):
            self.globalVariables[name] = 0

    def addSteps(self, integrator, fraction=1.0, force='f'):
        kT = _md(unit.BOLTZMANN_CONSTANT_kB*unit.AVOGADRO_CONSTANT_NA*self.temperature)
        NkT = self.degreesOfFreedom*kT
        tau = _md(self.timeScale)
        Q1, Q2 = NkT*tau**2, kT*tau**2
        half = 0.5*fraction
        kick2 = 'p_NHC_2 + (p_NHC_1^2/{}-{})*{}*dt'.format(Q1, kT, half)
        scale1 = 'p_NHC_1*exp(-{}*p_NHC_2*dt)'.format(half/Q2)
        integrator.addComputeGlobal('p_NHC_2', kick2)
        integrator.addComputeGlobal('p_NHC_1', scale1)
        integrator.addComputeGlobal('p_NHC_1', 'p_NHC_1 + (mvv-{})*{}*dt'.format(NkT, half))
        integrator.addComputeGlobal('vscaling', 'exp(-{}*p_NHC_1*dt)'.format(fraction/Q1))
        integrator.addComputeGlobal('p_NHC_1', 'p_NHC_1 + (vscaling^2*mvv-{})*{}*dt'.format(NkT, half))
        integrator.addComputeGlobal('p_NHC_1', scale1)
        integrator.addComputeGlobal('p_NHC_2', kick2)
        integrator.addComputePerDof('v', 'vscaling*v')


class NoseHooverLangevinPropagator(Propagator):
    """Nose-Hoover thermostat whose momentum is itself Langevin-thermostatted, B S O S B
    splitting (propagators.py:1452-1534)."""

    def __init__(self, temperature, degreesOfFreedom, timeScale, frictionConstant=None):
        super().__init__()
        self.temperature, self.degreesOfFreedom, self.timeScale = temperature, degreesOfFreedom, timeScale
        self.frictionConstant = 1/timeScale if frictionConstant is None else frictionConstant
        self.globalVariables['vscaling'] = 0
        self.globalVariables['p_NHL'] = 0

    def addSteps(self, integrator, fraction=1.0, force='f'):
        kT = _md(unit.BOLTZMANN_CONSTANT_kB*unit.AVOGADRO_CONSTANT_NA*self.temperature)
        NkT = self.degreesOfFreedom*kT
        Q = NkT*_md(self.timeScale)**2
        gamma = _md(self.frictionConstant)
        half = 0.5*fraction
        integrator.addComputeGlobal('p_NHL', 'p_NHL + (mvv-{})*{}*dt'.format(NkT, half))
        integrator.addComputeGlobal('vscaling', 'exp(-{}*p_NHL*dt)'.format(half/Q))
        integrator.addComputeGlobal('p_NHL', 'p_NHL*x + sqrt({}*(1-x^2))*gaussian; x = exp(-{}*dt)'.format(
            kT/Q, gamma*fraction))
        integrator.addComputeGlobal('vscaling', 'vscaling*exp(-{}*p_NHL*dt)'.format(half/Q))
        integrator.addComputeGlobal('p_NHL', 'p_NHL + (vscaling^2*mvv-{})*{}*dt'.format(NkT, half))
        integrator.addComputePerDof('v', 'vscaling*v')


class ExtendedSystemPropagator(Propagator):
    """Re-targets a per-DOF propagator onto one extended (AFED) variable
    (propagators.py:2119-2172): x -> parameter, v/m/... -> v_parameter/m_parameter/...,
    f -> -dE/dparameter, with periodic wrapping of the parameter into [-L/2, L/2]."""

    def __init__(self, parameter, mass, period, propagator, group=None):
        super().__init__()
        self._propagator = propagator
        self._parameter = parameter
        self._per_parameter_variables = ['v', 'm', 'L']
        self.globalVariables['v_' + parameter] = 0
        self.globalVariables['m_' + parameter] = mass
        self.globalVariables['L_' + parameter] = period
        for source in (propagator.perDofVariables, propagator.globalVariables):
            for name, value in source.items():
                self._per_parameter_variables.append(name)
                self.globalVariables['{}_{}'.format(name, parameter)] = value

    def _translate(self, expression):
        p = self._parameter
        out = re.sub(r'\bx\b', p, expression)
        for symbol in self._per_parameter_variables:
            out = re.sub(r'\b{}\b'.format(symbol), '{}_{}'.format(symbol, p), out)
        return re.sub(r'\bf([0-9]*)\b', '(-deriv(energy\\1,{}))'.format(p), out)

    def addSteps(self, integrator, fraction=1.0, force='f'):
        from . import integrators
        CI = mm.CustomIntegrator
        scratch = integrators._AtomsMM_Integrator(0)
        self._propagator.addSteps(scratch, fraction, force)
        p = self._parameter
        for k in range(scratch.getNumComputations()):
            kind, variable, expression = scratch.getComputationStep(k)
            target, translated = variable, expression
            if variable == 'x':
                target = p
                translated = 'select(step(-L/2-y),y+L,select(step(y-L/2),y-L,y)); L=L_{}; y={}'.format(
                    p, self._translate(expression))
            elif variable in self._per_parameter_variables:
                target = '{}_{}'.format(variable, p)
                translated = self._translate(expression)
            if kind in (CI.ComputeGlobal, CI.ComputePerDof):
                integrator.addComputeGlobal(target, translated)
            elif kind == CI.ComputeSum:
                raise Exception('ComputeSum not allowed in per-parameter integration')
            elif kind == CI.ConstrainPositions:
                integrator.addConstrainPositions()
            elif kind == CI.ConstrainVelocities:
                integrator.addConstrainVelocities()
            elif kind == CI.UpdateContextState:
                integrator.addUpdateContextState()
            elif kind == CI.IfBlockStart:
                integrator.beginIfBlock(translated)
            elif kind == CI.WhileBlockStart:
                integrator.beginWhileBlock(translated)
            elif kind == CI.BlockEnd:
                integrator.endBlock()

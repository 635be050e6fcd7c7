"""
Integrator descriptions: a step-program recorder with atomsmm's bookkeeping, the canned RESPA
integrators and the AFED (adiabatic free-energy dynamics) outer integrator.

Same names and behaviour as the reference's ``atomsmm.integrators`` (reference:
src/atomsmm/integrators.py:26-352, 642-860; rows a18-a20 of SURVEY 8a).  The recorded program
is executed by the CUDA engine (``engine.Context`` -> ``lowering.lower_program`` ->
csrc/integrate.cu); nothing here runs physics.

Unlike the reference, requirement analysis uses this package's own expression parser instead of
sympy, which (sympy >= 1.6) mis-parses the global name ``Q``.
"""

import functools
import math
import re
import types

import numpy as np

from . import expr as _expr
from . import mm
from . import propagators
from . import unit
from .unit import md_value as _md
from .utils import InputError
from .utils import kB

_STEP_FORMATS = ('{target} <- {expr}', '{target} <- {expr}', '{target} <- sum({expr})', 'constrain positions',
                 'constrain velocities', 'allow forces to update the context state', 'if ({expr}):',
                 'while ({expr}):', 'end')
_FORCE_NAME = re.compile(r'^f[0-9]*$')


class _AtomsMM_Integrator(mm.CustomIntegrator):
    """CustomIntegrator recorder that (i) keeps ``mvv = sum(m v v)`` up to date lazily, (ii)
    inserts one ``UpdateContextState`` before the first use of a force and (iii) routes
    expressions mixing several force groups through per-DOF buffers ``_fK_``
    (integrators.py:26-163)."""

    def __init__(self, stepSize):
        super().__init__(stepSize)
        self.addGlobalVariable('mvv', 0.0)
        self.addGlobalVariable('NDOF', 0.0)
        self.addPerDofVariable('ndof', 0.0)
        self._obsoleteKinetic = True
        self._obsoleteContextState = True
        self._random = np.random.RandomState()
        self._uninitialized = True

    def __repr__(self):
        lines = ['Per-dof variables:',
                 '  ' + ', '.join(self.getPerDofVariableName(k) for k in range(self.getNumPerDofVariables())),
                 'Global variables:']
        for k in range(self.getNumGlobalVariables()):
            lines.append('  {} = {}'.format(self.getGlobalVariableName(k), self.getGlobalVariable(k)))
        lines.append('Computation steps:')
        depth = 0
        for k in range(self.getNumComputations()):
            kind, target, expression = self.getComputationStep(k)
            if kind == self.BlockEnd:
                depth -= 1
            lines.append('{:4d}: '.format(k) + '   '*depth + _STEP_FORMATS[kind].format(target=target, expr=expression))
            if kind in (self.IfBlockStart, self.WhileBlockStart):
                depth += 1
        return '\n'.join(lines)

    def _normalVec(self):
        return mm.Vec3(self._random.normal(), self._random.normal(), self._random.normal())

    def _required_variables(self, variable, expression):
        return _expr.required_variables(variable, expression)

    def _checkUpdate(self, requirements):
        if self._obsoleteKinetic and 'mvv' in requirements:
            super().addComputeSum('mvv', 'm*v*v')
            self._obsoleteKinetic = False
        if self._obsoleteContextState and any(_FORCE_NAME.match(s) for s in requirements):
            super().addUpdateContextState()
            self._obsoleteContextState = False

    def addUpdateContextState(self):
        if self._obsoleteContextState:
            super().addUpdateContextState()
            self._obsoleteContextState = False

    def addComputeGlobal(self, variable, expression):
        if variable == 'mvv':
            raise InputError('Cannot assign value to global variable mvv')
        self._checkUpdate(self._required_variables(variable, expression))
        super().addComputeGlobal(variable, expression)

    def addComputePerDof(self, variable, expression):
        requirements = self._required_variables(variable, expression)
        self._checkUpdate(requirements)
        forces = sorted(s for s in requirements if _FORCE_NAME.match(s))
        if len(forces) > 1:
            # one force group per expression: all but the first go through per-DOF buffers
            expression = re.sub(r'\bf([0-9]*)\b', '_f\\1_', expression)
            known = [self.getPerDofVariableName(k) for k in range(self.getNumPerDofVariables())]
            for force in forces[1:]:
                buffer = '_{}_'.format(force)
                if buffer not in known:
                    self.addPerDofVariable(buffer, 0.0)
                self.addComputePerDof(buffer, force)
            expression = re.sub(r'\b_{}_\b'.format(forces[0]), forces[0], expression)
        super().addComputePerDof(variable, expression)
        if variable == 'v':
            self._obsoleteKinetic = True

    def setRandomNumberSeed(self, seed):
        self._random.seed(seed)
        super().setRandomNumberSeed(int(self._random.tomaxint() % 2**31))

    def step(self, steps):
        if self._uninitialized:
            if self._context is None:
                raise mm.OpenMMException('This Integrator is not bound to a context!')
            n = self._context._n
            self._ndof = ndof = 3*n
            self.setGlobalVariableByName('NDOF', ndof)
            self.setPerDofVariableByName('ndof', np.full((n, 3), float(ndof)))
            self.initialize()
            self._uninitialized = False
        return super().step(steps)

    def initialize(self):
        """Hook: random initialisation of thermostat variables before the first step."""
        pass


class GlobalThermostatIntegrator(_AtomsMM_Integrator):
    """``T(dt/2) NVE(dt) T(dt/2)`` with a global thermostat propagator T (integrators.py:173-211)."""

    def __init__(self, stepSize, nveIntegrator, thermostat=None):
        super().__init__(stepSize)
        whole = nveIntegrator if thermostat is None else \
            propagators.TrotterSuzukiPropagator(nveIntegrator, thermostat)
        whole.addVariables(self)
        whole.addSteps(self)


class MultipleTimeScaleIntegrator(_AtomsMM_Integrator):
    """RESPA integrator; arguments as :class:`propagators.MultipleTimeScalePropagator`
    (integrators.py:214-269)."""

    def __init__(self, stepSize, loops, move=None, boost=None, bath=None, **kwargs):
        super().__init__(stepSize)
        whole = propagators.MultipleTimeScalePropagator(loops, move, boost, bath, **kwargs)
        whole.addVariables(self)
        whole.addSteps(self)


class NHL_R_Integrator(MultipleTimeScaleIntegrator):
    """Massive Nose-Hoover-Langevin RESPA integrator: per-DOF thermostat velocity ``v2`` with
    Q2 = kT tau^2 (integrators.py:272-319)."""

    def __init__(self, stepSize, loops, temperature, timeScale, frictionConstant, **kwargs):
        scaling = propagators.GenericScalingPropagator('v', 'v2')
        ou = propagators.OrnsteinUhlenbeckPropagator(temperature, frictionConstant, 'v2', 'Q2', 'm*v^2 - kT',
                                                     Q2=kB*temperature*timeScale**2, kT=kB*temperature)
        super().__init__(stepSize, loops, None, None, propagators.TrotterSuzukiPropagator(ou, scaling), **kwargs)

    def initialize(self):
        kT = self.getGlobalVariableByName('kT')
        Q2 = self.getGlobalVariableByName('Q2')
        n = self._context._n
        self.setPerDofVariableByName('v2', math.sqrt(kT/Q2)*self._random.normal(size=(n, 3)))


class Langevin_R_Integrator(MultipleTimeScaleIntegrator):
    """Multiple-time-scale Langevin integrator: exact OU velocity update in the middle of the
    innermost drift (integrators.py:322-352)."""

    def __init__(self, stepSize, loops, temperature, frictionConstant, **kwargs):
        bath = propagators.OrnsteinUhlenbeckPropagator(temperature, frictionConstant, 'v', 'm', kT=kB*temperature)
        super().__init__(stepSize, loops, None, None, bath, **kwargs)


class ExtendedSystemVariable(object):
    """An AFED extended variable: a global context parameter ``name`` given a mass, its own
    temperature and a Nose-Hoover or Langevin thermostat, confined to
    [lower_limit, upper_limit] by reflecting walls or periodicity (integrators.py:642-744)."""

    def __init__(self, name, mass, kT, time_scale, lower_limit=0, upper_limit=1, periodic=False,
                 thermostat='Nose-Hoover', friction_constant=0.1/unit.femtoseconds):
        self._m_value = mass
        self._kT_value = kT
        self._Q_eta_value = kT*time_scale**2
        self._lower_limit, self._upper_limit, self._periodic = lower_limit, upper_limit, periodic
        self._gamma_value = friction_constant
        self._thermostat = thermostat
        self._x = name
        for tag in ('v', 'm', 'kT', 'kTbym', 'v_eta', 'Q_eta', 'gamma'):
            setattr(self, '_' + tag, '_{}_{}'.format(tag, name))

    def add_global_variables(self, integrator):
        integrator.addGlobalVariable(self._v, 0.0)
        integrator.addGlobalVariable(self._m, self._m_value)
        if self._thermostat == 'Nose-Hoover':
            integrator.addGlobalVariable(self._v_eta, 0.0)
            integrator.addGlobalVariable(self._kT, self._kT_value)
            integrator.addGlobalVariable(self._Q_eta, self._Q_eta_value)
        elif self._thermostat == 'Langevin':
            integrator.addGlobalVariable(self._kTbym, self._kT_value/self._m_value)
            integrator.addGlobalVariable(self._gamma, self._gamma_value)

    def _apply_boundary_conditions(self, integrator):
        above = 'step({}-({}))'.format(self._x, self._lower_limit)
        below = 'step({}-{})'.format(self._upper_limit, self._x)
        integrator.beginIfBlock('{}*{} = 0'.format(above, below))
        if self._periodic:
            length = self._upper_limit - self._lower_limit
            integrator.addComputeGlobal(self._x, '{} + select({},{},{})'.format(self._x, above, -length, length))
        else:
            integrator.addComputeGlobal(self._x, 'select({},{},{})-{}'.format(
                above, 2*self._upper_limit, 2*self._lower_limit, self._x))
            integrator.addComputeGlobal(self._v, '-{}'.format(self._v))
        integrator.endBlock()

    def add_integration_steps(self, integrator):
        move = '{} + 0.5*dt*{}'.format(self._x, self._v)
        integrator.addComputeGlobal(self._x, move)
        self._apply_boundary_conditions(integrator)
        if self._thermostat == 'Nose-Hoover':
            kick = '{0} + 0.5*dt*({1}*{2}^2-{3})/{4}'.format(self._v_eta, self._m, self._v, self._kT, self._Q_eta)
            integrator.addComputeGlobal(self._v_eta, kick)
            integrator.addComputeGlobal(self._v, '{}*exp(-dt*{})'.format(self._v, self._v_eta))
            integrator.addComputeGlobal(self._v_eta, kick)
        elif self._thermostat == 'Langevin':
            integrator.addComputeGlobal(self._v, 'z*{}+sqrt((1-z*z)*{})*gaussian; z=exp(-dt*{})'.format(
                self._v, self._kTbym, self._gamma))
        integrator.addComputeGlobal(self._x, move)
        self._apply_boundary_conditions(integrator)

    def update_velocity(self, integrator, divisor):
        integrator.addComputeGlobal(self._v, '{} - 0.5*(dt/{})*deriv(energy,{})/{}'.format(
            self._v, divisor, self._x, self._m))

    def initialize(self, integrator):
        sigma_v = math.sqrt(_md(self._kT_value)/_md(self._m_value))
        integrator.setGlobalVariableByName(self._v, sigma_v*integrator._random.normal())
        if self._thermostat == 'Nose-Hoover':
            sigma_eta = math.sqrt(_md(self._kT_value)/_md(self._Q_eta_value))
            integrator.setGlobalVariableByName(self._v_eta, sigma_eta*integrator._random.normal())


class AdiabaticDynamicsIntegrator(_AtomsMM_Integrator):
    """AFED outer integrator (integrators.py:747-860).  One step of size 2*nsteps*dt_inner is

        [kick_lambda  INNER(dt_inner)  kick_lambda]^nsteps   move/thermostat lambda   [same]^nsteps

    where INNER is the program of ``custom_integrator`` with ``dt`` rewritten to
    ``dt/(2*nsteps)`` and kick_lambda uses ``-deriv(energy, lambda)``.
    """

    def __init__(self, custom_integrator, nsteps, variables):
        super().__init__(2*nsteps*custom_integrator.getStepSize())
        self._variables = variables
        if nsteps > 1:
            self._counter = '_nsteps_counter'
            self.addGlobalVariable(self._counter, 0)
        for variable in variables:
            variable.add_global_variables(self)
        self._import_variables_and_initializer(custom_integrator)
        self.addUpdateContextState()
        self._add_physical_steps(custom_integrator, nsteps)
        for variable in variables:
            variable.add_integration_steps(self)
        self._add_physical_steps(custom_integrator, nsteps)

    def _add_physical_steps(self, integrator, nsteps):
        if nsteps > 1:
            self.addComputeGlobal(self._counter, '0')
            self.beginWhileBlock('{} < {}'.format(self._counter, nsteps))
        for variable in self._variables:
            variable.update_velocity(self, 2*nsteps)
        self._import_computations(integrator, nsteps)
        for variable in self._variables:
            variable.update_velocity(self, 2*nsteps)
        if nsteps > 1:
            self.addComputeGlobal(self._counter, '{} + 1'.format(self._counter))
            self.endBlock()

    def _import_computations(self, integrator, nsteps):
        dt = re.compile(r'\bdt\b')
        for k in range(integrator.getNumComputations()):
            kind, variable, expression = integrator.getComputationStep(k)
            expression = dt.sub('(dt/{})'.format(2*nsteps), expression)
            if kind == self.ComputeGlobal:
                self.addComputeGlobal(variable, expression)
            elif kind == self.ComputePerDof:
                self.addComputePerDof(variable, expression)
            elif kind == self.ComputeSum:
                self.addComputeSum(variable, expression)
            elif kind == self.ConstrainPositions:
                self.addConstrainPositions()
            elif kind == self.ConstrainVelocities:
                self.addConstrainVelocities()
            elif kind == self.UpdateContextState:
                self.addUpdateContextState()
            elif kind == self.IfBlockStart:
                self.beginIfBlock(expression)
            elif kind == self.WhileBlockStart:
                self.beginWhileBlock(expression)
            elif kind == self.BlockEnd:
                self.endBlock()

    def _import_variables_and_initializer(self, integrator):
        for k in range(integrator.getNumGlobalVariables()):
            name = integrator.getGlobalVariableName(k)
            if name not in ('mvv', 'NDOF'):
                self.addGlobalVariable(name, integrator.getGlobalVariable(k))
        for k in range(integrator.getNumPerDofVariables()):
            name = integrator.getPerDofVariableName(k)
            if name != 'ndof':
                self.addPerDofVariable(name, 0)
                value = integrator._perdof_values[k]
                if not np.isscalar(value):
                    self.setPerDofVariableByName(name, value)
        source = integrator.initialize
        function = getattr(source, '__func__', source)
        clone = types.FunctionType(function.__code__, function.__globals__)
        self._initialize_function = functools.update_wrapper(clone, function)

    def initialize(self):
        self._initialize_function(self)
        for variable in self._variables:
            variable.initialize(self)

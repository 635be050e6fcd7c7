"""
Description layer: the part of OpenMM's public object model that atomsmm drives.

atomsmm never computes anything itself; it *describes* systems (``System`` + ``Force``
objects carrying parameters and algebraic energy strings) and integrators
(``CustomIntegrator`` step programs) and hands them to OpenMM (call-site
census from ``grep`` over reference ``src/atomsmm``: addComputePerDof, addComputeGlobal,
setForceGroup, addForce, addGlobalParameter, getState, getParticleParameters, ...).
This module re-provides exactly those description classes as plain Python containers, with
the same method names, argument meaning and unit conventions, so that the atomsmm classes
in this package read like the reference's.  Nothing here evaluates physics: evaluation is
done by ``engine.Context`` over the CUDA C-ABI (csrc/), and only there.

All values are stored as floats in MD units (nm, ps, dalton, kJ/mol, e, K, rad); getters
return ``unit.Quantity`` objects like OpenMM's do.
"""

import copy

import numpy as np

from . import unit
from .unit import md_value as _md


class OpenMMException(Exception):
    pass


class Vec3(tuple):
    def __new__(cls, x, y, z):
        return tuple.__new__(cls, (x, y, z))

    def __getnewargs__(self):
        return tuple(self)

    x = property(lambda self: self[0])
    y = property(lambda self: self[1])
    z = property(lambda self: self[2])

    def __add__(self, other):
        return Vec3(self[0] + other[0], self[1] + other[1], self[2] + other[2])

    def __sub__(self, other):
        return Vec3(self[0] - other[0], self[1] - other[1], self[2] - other[2])

    def __mul__(self, other):
        if isinstance(other, unit.Unit):
            return unit.Quantity(self, other)
        return Vec3(self[0]*other, self[1]*other, self[2]*other)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return Vec3(self[0]/other, self[1]/other, self[2]/other)

    def __neg__(self):
        return Vec3(-self[0], -self[1], -self[2])


class PackedRows(object):
    """Columnar, read-mostly storage of per-particle / per-term records: ``indices`` (int32 [n, k],
    k may be 0) followed by ``values`` (float64 [n, m]).  Behaves like the list of rows the
    description classes keep (``len``, indexing, iteration, item assignment), so getters work
    unchanged, while consumers that need whole tables (engine.Context, replication of a box to
    millions of atoms) take the arrays directly through ``index_columns`` / ``value_columns``.
    ``nested``: the values of a row form ONE trailing tuple (CustomBondForce / CustomAngleForce rows)."""

    def __init__(self, indices, values, nested=False):
        n = len(values) if values is not None else len(indices)
        self.indices = np.ascontiguousarray(indices if indices is not None else np.zeros((n, 0)), dtype=np.int32)
        self.values = np.ascontiguousarray(values if values is not None else np.zeros((n, 0)), dtype=np.float64)
        self.indices = self.indices.reshape(n, -1)
        self.values = self.values.reshape(n, -1)
        self.nested = bool(nested)

    def __len__(self):
        return self.values.shape[0]

    def _row(self, k):
        head = [int(i) for i in self.indices[k]]
        tail = [float(v) for v in self.values[k]]
        return head + ([tail] if self.nested else tail)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self._row(j) for j in range(*k.indices(len(self)))]
        return self._row(k)

    def __setitem__(self, k, row):
        ni = self.indices.shape[1]
        self.indices[k] = [int(i) for i in row[:ni]]
        tail = row[ni] if self.nested else row[ni:]
        self.values[k] = [float(v) for v in tail]

    def __iter__(self):
        for k in range(len(self)):
            yield self._row(k)

    def append(self, row):
        raise OpenMMException('this force holds a packed (replicated) table: rows cannot be appended')


def index_columns(rows, count):
    """int32 [n, count]: the leading index columns of a list of rows or a PackedRows."""
    if isinstance(rows, PackedRows):
        return rows.indices[:, :count]
    return np.array([r[:count] for r in rows], dtype=np.int32).reshape(len(rows), count)


def value_columns(rows, first):
    """float64 [n, m]: the values after ``first`` index columns (a nested trailing tuple is flattened)."""
    if isinstance(rows, PackedRows):
        return rows.values
    if len(rows) == 0:
        return np.zeros((0, 0), dtype=np.float64)
    if len(rows[0]) == first + 1 and isinstance(rows[0][first], (list, tuple)):
        return np.array([r[first] for r in rows], dtype=np.float64).reshape(len(rows), -1)
    return np.array([r[first:] for r in rows], dtype=np.float64).reshape(len(rows), -1)



_nm = unit.nanometer
_kj = unit.kilojoule_per_mole
_e = unit.elementary_charge


class Force(object):
    def __init__(self):
        self._group = 0
        self._periodic = False

    def getForceGroup(self):
        return self._group

    def setForceGroup(self, group):
        if not 0 <= int(group) <= 31:
            raise OpenMMException('Force group must be between 0 and 31')
        self._group = int(group)

    def usesPeriodicBoundaryConditions(self):
        return self._periodic

    def setUsesPeriodicBoundaryConditions(self, periodic):
        self._periodic = bool(periodic)

    def updateParametersInContext(self, context):
        context._parameters_changed(self)

    def __copy__(self):
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        return new


class _GlobalParameterMixin(object):
    def _init_globals(self):
        self._globals = []          # [name, default]
        self._derivs = []

    def addGlobalParameter(self, name, value):
        self._globals.append([name, float(_md(value))])
        return len(self._globals) - 1

    def getNumGlobalParameters(self):
        return len(self._globals)

    def getGlobalParameterName(self, index):
        return self._globals[index][0]

    def getGlobalParameterDefaultValue(self, index):
        return self._globals[index][1]

    def setGlobalParameterDefaultValue(self, index, value):
        self._globals[index][1] = float(_md(value))

    def addEnergyParameterDerivative(self, name):
        if name not in self._derivs:
            self._derivs.append(name)

    def getNumEnergyParameterDerivatives(self):
        return len(self._derivs)

    def getEnergyParameterDerivativeName(self, index):
        return self._derivs[index]


class NonbondedForce(Force, _GlobalParameterMixin):
    NoCutoff, CutoffNonPeriodic, CutoffPeriodic, Ewald, PME, LJPME = range(6)

    def __init__(self):
        Force.__init__(self)
        self._init_globals()
        self._particles = []       # [q, sigma, eps]
        self._exceptions = []      # [i, j, qq, sigma, eps]
        self._exception_index = {}
        self._method = self.NoCutoff
        self._cutoff = 1.0
        self._use_switch = False
        self._switch = -1.0
        self._rf_dielectric = 78.3
        self._ewald_tol = 5e-4
        self._pme = (0.0, 0, 0, 0)
        self._use_lrc = True
        self._recip_group = -1
        self._particle_offsets = []
        self._exception_offsets = []

    def usesPeriodicBoundaryConditions(self):
        return self._method in (self.CutoffPeriodic, self.Ewald, self.PME, self.LJPME)

    def getNumParticles(self):
        return len(self._particles)

    def getNumExceptions(self):
        return len(self._exceptions)

    def addParticle(self, charge, sigma, epsilon):
        self._particles.append([float(_md(charge)), float(_md(sigma)), float(_md(epsilon))])
        return len(self._particles) - 1

    def getParticleParameters(self, index):
        q, s, e = self._particles[index]
        return [q*_e, s*_nm, e*_kj]

    def setParticleParameters(self, index, charge, sigma, epsilon):
        self._particles[index] = [float(_md(charge)), float(_md(sigma)), float(_md(epsilon))]

    def addException(self, particle1, particle2, chargeProd, sigma, epsilon, replace=False):
        key = (min(particle1, particle2), max(particle1, particle2))
        entry = [int(particle1), int(particle2), float(_md(chargeProd)), float(_md(sigma)), float(_md(epsilon))]
        if key in self._exception_index:
            if not replace:
                raise OpenMMException('NonbondedForce: There is already an exception for particles %d and %d' % key)
            self._exceptions[self._exception_index[key]] = entry
            return self._exception_index[key]
        self._exception_index[key] = len(self._exceptions)
        self._exceptions.append(entry)
        return len(self._exceptions) - 1

    def getExceptionParameters(self, index):
        i, j, qq, s, e = self._exceptions[index]
        return [i, j, qq*_e**2, s*_nm, e*_kj]

    def setExceptionParameters(self, index, particle1, particle2, chargeProd, sigma, epsilon):
        self._exceptions[index] = [int(particle1), int(particle2), float(_md(chargeProd)),
                                   float(_md(sigma)), float(_md(epsilon))]

    def createExceptionsFromBonds(self, bonds, coulomb14Scale, lj14Scale):
        """1-2 and 1-3 pairs become exclusions, 1-4 pairs scaled exceptions."""
        n = self.getNumParticles()
        bonded = [set() for _ in range(n)]
        for a, b in bonds:
            bonded[a].add(b)
            bonded[b].add(a)
        excl12_13 = set()
        pairs14 = set()
        for a in range(n):
            for b in bonded[a]:
                excl12_13.add((min(a, b), max(a, b)))
                for c in bonded[b]:
                    if c == a:
                        continue
                    excl12_13.add((min(a, c), max(a, c)))
                    for d in bonded[c]:
                        if d != b and d != a:
                            pairs14.add((min(a, d), max(a, d)))
        for (a, d) in sorted(pairs14 - excl12_13):
            qa, sa, ea = self._particles[a]
            qd, sd, ed = self._particles[d]
            self.addException(a, d, coulomb14Scale*qa*qd, 0.5*(sa + sd), lj14Scale*np.sqrt(ea*ed))
        for (a, b) in sorted(excl12_13):
            self.addException(a, b, 0.0, 1.0, 0.0)

    def getNonbondedMethod(self):
        return self._method

    def setNonbondedMethod(self, method):
        self._method = int(method)

    def getCutoffDistance(self):
        return self._cutoff*_nm

    def setCutoffDistance(self, distance):
        self._cutoff = float(_md(distance))

    def getUseSwitchingFunction(self):
        return self._use_switch

    def setUseSwitchingFunction(self, use):
        self._use_switch = bool(use)

    def getSwitchingDistance(self):
        return self._switch*_nm

    def setSwitchingDistance(self, distance):
        self._switch = float(_md(distance))

    def getReactionFieldDielectric(self):
        return self._rf_dielectric

    def setReactionFieldDielectric(self, dielectric):
        self._rf_dielectric = float(dielectric)

    def getEwaldErrorTolerance(self):
        return self._ewald_tol

    def setEwaldErrorTolerance(self, tol):
        self._ewald_tol = float(tol)

    def getPMEParameters(self):
        alpha, nx, ny, nz = self._pme
        return [alpha/_nm, nx, ny, nz]

    def setPMEParameters(self, alpha, nx, ny, nz):
        self._pme = (float(_md(alpha)), int(nx), int(ny), int(nz))

    def getUseDispersionCorrection(self):
        return self._use_lrc

    def setUseDispersionCorrection(self, use):
        self._use_lrc = bool(use)

    def getReciprocalSpaceForceGroup(self):
        return self._recip_group

    def setReciprocalSpaceForceGroup(self, group):
        self._recip_group = int(group)

    # parameter offsets (alchemical systems; carried, not evaluated by the hot path)
    def addParticleParameterOffset(self, parameter, particleIndex, chargeScale, sigmaScale, epsilonScale):
        self._particle_offsets.append([parameter, int(particleIndex), float(_md(chargeScale)),
                                       float(_md(sigmaScale)), float(_md(epsilonScale))])
        return len(self._particle_offsets) - 1

    def getNumParticleParameterOffsets(self):
        return len(self._particle_offsets)

    def getParticleParameterOffset(self, index):
        return list(self._particle_offsets[index])

    def addExceptionParameterOffset(self, parameter, exceptionIndex, chargeProdScale, sigmaScale, epsilonScale):
        self._exception_offsets.append([parameter, int(exceptionIndex), float(_md(chargeProdScale)),
                                        float(_md(sigmaScale)), float(_md(epsilonScale))])
        return len(self._exception_offsets) - 1

    def getNumExceptionParameterOffsets(self):
        return len(self._exception_offsets)

    def getExceptionParameterOffset(self, index):
        return list(self._exception_offsets[index])


class _CustomParametrized(Force, _GlobalParameterMixin):
    def __init__(self, energy):
        Force.__init__(self)
        self._init_globals()
        self._energy = str(energy)

    def getEnergyFunction(self):
        return self._energy

    def setEnergyFunction(self, energy):
        self._energy = str(energy)


class CustomNonbondedForce(_CustomParametrized):
    NoCutoff, CutoffNonPeriodic, CutoffPeriodic = range(3)

    def __init__(self, energy):
        _CustomParametrized.__init__(self, energy)
        self._per_particle = []
        self._particles = []
        self._exclusions = []
        self._method = self.NoCutoff
        self._cutoff = 1.0
        self._use_switch = False
        self._switch = -1.0
        self._use_lrc = False
        self._interaction_groups = []

    def usesPeriodicBoundaryConditions(self):
        return self._method == self.CutoffPeriodic

    def addPerParticleParameter(self, name):
        self._per_particle.append(name)
        return len(self._per_particle) - 1

    def getNumPerParticleParameters(self):
        return len(self._per_particle)

    def getPerParticleParameterName(self, index):
        return self._per_particle[index]

    def addParticle(self, parameters=()):
        self._particles.append([float(_md(p)) for p in parameters])
        return len(self._particles) - 1

    def getNumParticles(self):
        return len(self._particles)

    def getParticleParameters(self, index):
        return tuple(self._particles[index])

    def setParticleParameters(self, index, parameters):
        self._particles[index] = [float(_md(p)) for p in parameters]

    def addExclusion(self, particle1, particle2):
        self._exclusions.append((int(particle1), int(particle2)))
        return len(self._exclusions) - 1

    def getNumExclusions(self):
        return len(self._exclusions)

    def getExclusionParticles(self, index):
        return list(self._exclusions[index])

    def getNonbondedMethod(self):
        return self._method

    def setNonbondedMethod(self, method):
        self._method = int(method)

    def getCutoffDistance(self):
        return self._cutoff*_nm

    def setCutoffDistance(self, distance):
        self._cutoff = float(_md(distance))

    def getUseSwitchingFunction(self):
        return self._use_switch

    def setUseSwitchingFunction(self, use):
        self._use_switch = bool(use)

    def getSwitchingDistance(self):
        return self._switch*_nm

    def setSwitchingDistance(self, distance):
        self._switch = float(_md(distance))

    def getUseLongRangeCorrection(self):
        return self._use_lrc

    def setUseLongRangeCorrection(self, use):
        self._use_lrc = bool(use)

    def addInteractionGroup(self, set1, set2):
        self._interaction_groups.append((set(set1), set(set2)))
        return len(self._interaction_groups) - 1

    def getNumInteractionGroups(self):
        return len(self._interaction_groups)

    def getInteractionGroupParameters(self, index):
        return self._interaction_groups[index]


class CustomBondForce(_CustomParametrized):
    def __init__(self, energy):
        _CustomParametrized.__init__(self, energy)
        self._per_bond = []
        self._bonds = []

    def addPerBondParameter(self, name):
        self._per_bond.append(name)
        return len(self._per_bond) - 1

    def getNumPerBondParameters(self):
        return len(self._per_bond)

    def getPerBondParameterName(self, index):
        return self._per_bond[index]

    def addBond(self, particle1, particle2, parameters=()):
        self._bonds.append([int(particle1), int(particle2), [float(_md(p)) for p in parameters]])
        return len(self._bonds) - 1

    def getNumBonds(self):
        return len(self._bonds)

    def getBondParameters(self, index):
        i, j, p = self._bonds[index]
        return [i, j, tuple(p)]

    def setBondParameters(self, index, particle1, particle2, parameters=()):
        self._bonds[index] = [int(particle1), int(particle2), [float(_md(p)) for p in parameters]]


class CustomAngleForce(_CustomParametrized):
    def __init__(self, energy):
        _CustomParametrized.__init__(self, energy)
        self._per_angle = []
        self._angles = []

    def addPerAngleParameter(self, name):
        self._per_angle.append(name)
        return len(self._per_angle) - 1

    def getNumPerAngleParameters(self):
        return len(self._per_angle)

    def getPerAngleParameterName(self, index):
        return self._per_angle[index]

    def addAngle(self, particle1, particle2, particle3, parameters=()):
        self._angles.append([int(particle1), int(particle2), int(particle3), [float(_md(p)) for p in parameters]])
        return len(self._angles) - 1

    def getNumAngles(self):
        return len(self._angles)

    def getAngleParameters(self, index):
        i, j, k, p = self._angles[index]
        return [i, j, k, tuple(p)]


class HarmonicBondForce(Force):
    def __init__(self):
        Force.__init__(self)
        self._bonds = []

    def addBond(self, particle1, particle2, length, k):
        self._bonds.append([int(particle1), int(particle2), float(_md(length)), float(_md(k))])
        return len(self._bonds) - 1

    def getNumBonds(self):
        return len(self._bonds)

    def getBondParameters(self, index):
        i, j, r0, k = self._bonds[index]
        return [i, j, r0*_nm, k*_kj/_nm**2]

    def setBondParameters(self, index, particle1, particle2, length, k):
        self._bonds[index] = [int(particle1), int(particle2), float(_md(length)), float(_md(k))]


class HarmonicAngleForce(Force):
    def __init__(self):
        Force.__init__(self)
        self._angles = []

    def addAngle(self, particle1, particle2, particle3, angle, k):
        self._angles.append([int(particle1), int(particle2), int(particle3), float(_md(angle)), float(_md(k))])
        return len(self._angles) - 1

    def getNumAngles(self):
        return len(self._angles)

    def getAngleParameters(self, index):
        i, j, k, theta, K = self._angles[index]
        return [i, j, k, theta*unit.radian, K*_kj/unit.radian**2]

    def setAngleParameters(self, index, particle1, particle2, particle3, angle, k):
        self._angles[index] = [int(particle1), int(particle2), int(particle3), float(_md(angle)), float(_md(k))]


class PeriodicTorsionForce(Force):
    def __init__(self):
        Force.__init__(self)
        self._torsions = []

    def addTorsion(self, p1, p2, p3, p4, periodicity, phase, k):
        self._torsions.append([int(p1), int(p2), int(p3), int(p4), int(periodicity), float(_md(phase)), float(_md(k))])
        return len(self._torsions) - 1

    def getNumTorsions(self):
        return len(self._torsions)

    def getTorsionParameters(self, index):
        a, b, c, d, n, phase, k = self._torsions[index]
        return [a, b, c, d, n, phase*unit.radian, k*_kj]


class CMMotionRemover(Force):
    def __init__(self, frequency=1):
        Force.__init__(self)
        self._frequency = int(frequency)

    def getFrequency(self):
        return self._frequency


class MonteCarloBarostat(Force):
    """openmm.MonteCarloBarostat: isotropic Monte Carlo volume moves applied where the integrator calls
    UpdateContextState (the first computation of every atomsmm step program, integrators.py:115-122).
    The reference has no barostat of its own; this is the OpenMM force a user adds to
    the System for NPT runs."""

    def __init__(self, defaultPressure, defaultTemperature, frequency=25):
        Force.__init__(self)
        self._pressure = float(_md(defaultPressure.value_in_unit(unit.bar)
                                   if isinstance(defaultPressure, unit.Quantity) else defaultPressure))
        self._temperature = float(_md(defaultTemperature))
        self._frequency = int(frequency)
        self._seed = 0

    @staticmethod
    def Pressure():
        return 'MonteCarloPressure'

    @staticmethod
    def Temperature():
        return 'MonteCarloTemperature'

    def getDefaultPressure(self):
        return self._pressure*unit.bar

    def setDefaultPressure(self, pressure):
        self._pressure = float(pressure.value_in_unit(unit.bar) if isinstance(pressure, unit.Quantity) else pressure)

    def getDefaultTemperature(self):
        return self._temperature*unit.kelvin

    def setDefaultTemperature(self, temperature):
        self._temperature = float(_md(temperature))

    def getFrequency(self):
        return self._frequency

    def setFrequency(self, frequency):
        self._frequency = int(frequency)

    def getRandomNumberSeed(self):
        return self._seed

    def setRandomNumberSeed(self, seed):
        self._seed = int(seed)

    def usesPeriodicBoundaryConditions(self):
        return True


class System(object):
    def __init__(self):
        self._masses = []
        self._forces = []
        self._constraints = []
        self._box = [Vec3(2, 0, 0), Vec3(0, 2, 0), Vec3(0, 0, 2)]

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, copy.deepcopy(v, memo))
        return new

    # the reference clones systems with ``self.this = copy.deepcopy(system).this``
    # (systems.py:28,63); our equivalent is _adopt().
    def _adopt(self, system):
        clone = copy.deepcopy(system)
        self.__dict__.update(clone.__dict__)

    # SWIG-proxy idiom used by code written against OpenMM: ``self.this = deepcopy(system).this``
    @property
    def this(self):
        return self

    @this.setter
    def this(self, other):
        self.__dict__.update(copy.deepcopy(other).__dict__)

    def addParticle(self, mass):
        self._masses.append(float(_md(mass)))
        return len(self._masses) - 1

    def getNumParticles(self):
        return len(self._masses)

    def getParticleMass(self, index):
        return self._masses[index]*unit.dalton

    def setParticleMass(self, index, mass):
        self._masses[index] = float(_md(mass))

    def addConstraint(self, particle1, particle2, distance):
        self._constraints.append([int(particle1), int(particle2), float(_md(distance))])
        return len(self._constraints) - 1

    def getNumConstraints(self):
        return len(self._constraints)

    def getConstraintParameters(self, index):
        i, j, d = self._constraints[index]
        return [i, j, d*_nm]

    def addForce(self, force):
        self._forces.append(force)
        return len(self._forces) - 1

    def getNumForces(self):
        return len(self._forces)

    def getForce(self, index):
        return self._forces[index]

    def getForces(self):
        return list(self._forces)

    def removeForce(self, index):
        del self._forces[index]

    def getDefaultPeriodicBoxVectors(self):
        return [Vec3(*v)*_nm for v in self._box]

    def setDefaultPeriodicBoxVectors(self, a, b, c):
        self._box = [Vec3(*[float(x) for x in _md(v)]) for v in (a, b, c)]

    def usesPeriodicBoundaryConditions(self):
        return any(f.usesPeriodicBoundaryConditions() for f in self._forces)


# ---------------------------------------------------------------------------------------------
# Integrators (descriptions only)
# ---------------------------------------------------------------------------------------------

class Integrator(object):
    def __init__(self, stepSize):
        self._dt = float(_md(stepSize))
        self._context = None
        self._seed = 0
        self._constraint_tolerance = 1e-5

    def getConstraintTolerance(self):
        return self._constraint_tolerance

    def setConstraintTolerance(self, tolerance):
        self._constraint_tolerance = float(tolerance)
        if self._context is not None:
            self._context._integrator_changed()

    def getStepSize(self):
        return self._dt*unit.picosecond

    def setStepSize(self, stepSize):
        self._dt = float(_md(stepSize))
        if self._context is not None:
            self._context._integrator_changed()

    def getRandomNumberSeed(self):
        return self._seed

    def setRandomNumberSeed(self, seed):
        self._seed = int(seed)
        if self._context is not None:
            self._context._integrator_changed()

    def step(self, steps):
        if self._context is None:
            raise OpenMMException('This Integrator is not bound to a context!')
        self._context._step(int(steps))


class VerletIntegrator(Integrator):
    pass


class CustomIntegrator(Integrator):
    """Recorder for a step program; same step-kind codes as OpenMM (integrators.py:64-74)."""
    ComputeGlobal, ComputePerDof, ComputeSum, ConstrainPositions, ConstrainVelocities, \
        UpdateContextState, IfBlockStart, WhileBlockStart, BlockEnd = range(9)

    def __init__(self, stepSize):
        Integrator.__init__(self, stepSize)
        self._global_names = []
        self._global_values = []
        self._perdof_names = []
        self._perdof_values = []     # scalar default or ndarray [N,3]
        self._steps = []             # (kind, variable, expression)

    # variables ----------------------------------------------------------------------------
    def addGlobalVariable(self, name, initialValue):
        self._global_names.append(name)
        self._global_values.append(float(_md(initialValue)))
        return len(self._global_names) - 1

    def addPerDofVariable(self, name, initialValue):
        self._perdof_names.append(name)
        self._perdof_values.append(float(_md(initialValue)))
        return len(self._perdof_names) - 1

    def getNumGlobalVariables(self):
        return len(self._global_names)

    def getNumPerDofVariables(self):
        return len(self._perdof_names)

    def getGlobalVariableName(self, index):
        return self._global_names[index]

    def getPerDofVariableName(self, index):
        return self._perdof_names[index]

    def getGlobalVariable(self, index):
        if self._context is not None:
            return self._context._get_global(self._global_names[index])
        return self._global_values[index]

    def getGlobalVariableByName(self, name):
        return self.getGlobalVariable(self._global_names.index(name))

    def setGlobalVariable(self, index, value):
        self._global_values[index] = float(_md(value))
        if self._context is not None:
            self._context._set_global(self._global_names[index], self._global_values[index])

    def setGlobalVariableByName(self, name, value):
        self.setGlobalVariable(self._global_names.index(name), value)

    def getPerDofVariable(self, index):
        if self._context is not None:
            return self._context._get_perdof(self._perdof_names[index])
        value = self._perdof_values[index]
        if np.isscalar(value):
            n = 0 if self._context is None else self._context._n
            return [Vec3(value, value, value) for _ in range(n)]
        return [Vec3(*row) for row in value]

    def getPerDofVariableByName(self, name):
        return self.getPerDofVariable(self._perdof_names.index(name))

    def setPerDofVariable(self, index, values):
        array = np.asarray(_md(values), dtype=np.float64).reshape(-1, 3)
        self._perdof_values[index] = array
        if self._context is not None:
            self._context._set_perdof(self._perdof_names[index], array)

    def setPerDofVariableByName(self, name, values):
        self.setPerDofVariable(self._perdof_names.index(name), values)

    # steps --------------------------------------------------------------------------------
    def _add(self, kind, variable='', expression=''):
        if self._context is not None:
            raise OpenMMException('The integrator cannot be modified after it is bound to a context')
        self._steps.append((kind, variable, expression))
        return len(self._steps) - 1

    def addComputeGlobal(self, variable, expression):
        return self._add(self.ComputeGlobal, variable, expression)

    def addComputePerDof(self, variable, expression):
        return self._add(self.ComputePerDof, variable, expression)

    def addComputeSum(self, variable, expression):
        return self._add(self.ComputeSum, variable, expression)

    def addConstrainPositions(self):
        return self._add(self.ConstrainPositions)

    def addConstrainVelocities(self):
        return self._add(self.ConstrainVelocities)

    def addUpdateContextState(self):
        return self._add(self.UpdateContextState)

    def beginIfBlock(self, condition):
        return self._add(self.IfBlockStart, '', condition)

    def beginWhileBlock(self, condition):
        return self._add(self.WhileBlockStart, '', condition)

    def endBlock(self):
        return self._add(self.BlockEnd)

    def getNumComputations(self):
        return len(self._steps)

    def getComputationStep(self, index):
        return list(self._steps[index])


class Platform(object):
    """Only one platform exists: the sm_100a CUDA engine.  There is no CPU path."""
    _NAMES = ('B200', 'CUDA')

    def __init__(self, name):
        self._name = name
        self._properties = {}

    @staticmethod
    def getPlatformByName(name):
        if name not in Platform._NAMES:
            raise OpenMMException('There is no registered Platform called "%s": this engine has a single '
                                  'sm_100a CUDA platform ("B200"); there is deliberately no CPU fallback' % name)
        return Platform(name)

    @staticmethod
    def getNumPlatforms():
        return 1

    def getName(self):
        return self._name

    def setPropertyDefaultValue(self, name, value):
        self._properties[name] = value

    def getPropertyDefaultValue(self, name):
        return self._properties.get(name, '')


def __getattr__(name):
    # Context/State live in engine.py (they need the CUDA library); import lazily so that the
    # description layer alone works on machines without the built extension.
    if name in ('Context', 'State'):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)

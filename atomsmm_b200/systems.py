"""
System builders: ``RESPASystem`` (force-group layout for multiple-time-scale integration) and
``ComputingSystem`` (the virial-as-energies system behind ``PressureComputer``).

Same names, arguments and resulting force-group layout as the reference's
``atomsmm.systems`` (reference: src/atomsmm/systems.py:34-238, 868-944; rows a8/a9 of
SURVEY 8a).  ``AlchemicalSystem`` (systems.py:318-410) is provided with its default ``softcore``
coupling because it is the system of the AFED configuration (BASELINE config 4, row a20); the other
alchemical builders of the reference (SolvationSystem, AlchemicalRespaSystem and the
non-softcore couplings) are system-construction bookkeeping outside the hot path and are not provided.
"""

import copy
import re

from . import expr as _expr
from . import forces as _forces
from . import mm
from . import unit


class _AtomsMM_System(mm.System):
    def __init__(self, system, copyForces=True):
        super().__init__()
        self._adopt(system)
        if not copyForces:
            self._forces = []


class RESPASystem(mm.System):
    """A copy of ``system`` split for RESPA (systems.py:62-95):

    * group 2: the original NonbondedForce (direct and reciprocal space);
    * group 1: near potential (cutoff ``rcutIn``, switched from ``rswitchIn``) built from the
      same particles, all exceptions turned into exclusions;
    * group 31: minus the near potential, so that the sum of all groups is the original energy;
    * group 0: every bonded term plus (``fastExceptions``) the 1-4 exceptions, extracted from
      the NonbondedForce into an explicit-pair force.

    A ``RespaPropagator([n0, n1, 1])`` then kicks with f0, f1 and (f2 - f1).
    """

    def __init__(self, system, rcutIn, rswitchIn, **kwargs):
        super().__init__()
        self._adopt(system)
        self._special_bond_force = None
        self._special_angle_force = None
        adjustment = kwargs.pop('adjustment', 'force-switch')
        fast_exceptions = kwargs.get('fastExceptions', True)
        plain_ljc = ['4*epsilon*x*(x-1) + Kc*chargeprod/r', 'x=(sigma/r)^6', 'Kc=138.935456']
        for force in self.getForces():
            if not isinstance(force, mm.NonbondedForce):
                continue
            near = _forces.nearForceExpressions(rcutIn, rswitchIn, adjustment)
            minus_near = list(near)
            minus_near[0] = '-step(rc0-r)*({})'.format(near[0])
            force.setForceGroup(2)
            force.setReciprocalSpaceForceGroup(2)
            self._addCustomNonbondedForce(near, rcutIn, 1, force)
            self._addCustomNonbondedForce(minus_near, rcutIn, 31, force)
            if fast_exceptions:
                self._addCustomBondForce(plain_ljc, 0, force, extract=True)
            else:
                self._addCustomBondForce(near, 1, force)
                self._addCustomBondForce(minus_near, 31, force)

    def _addCustomNonbondedForce(self, expressions, rcut, group, source):
        force = _forces._AtomsMM_CustomNonbondedForce(';'.join(expressions), rcut, use_switching_function=False,
                                                      use_dispersion_correction=False)
        force.importFrom(source)
        force.setForceGroup(group)
        self.addForce(force)

    def _addCustomBondForce(self, expressions, group, nonbonded, extract=False):
        force = _forces._AtomsMM_CustomBondForce(';'.join(expressions))
        force.importFrom(nonbonded, extract)
        if force.getNumBonds() > 0:
            force.setForceGroup(group)
            self.addForce(force)

    @staticmethod
    def _matcher(topology, residue, names):
        resname = [atom.residue.name for atom in topology.atoms()]
        atomname = [atom.name for atom in topology.atoms()]
        r_regex = re.compile(residue)
        a_regex = [re.compile(n) for n in names]

        def match(*indices):
            if not all(r_regex.match(resname[j]) for j in indices):
                return False
            forward = all(rx.match(atomname[j]) for rx, j in zip(a_regex, indices))
            backward = all(rx.match(atomname[j]) for rx, j in zip(a_regex, indices[::-1]))
            return forward or backward
        return match

    def redefine_bond(self, topology, residue, atom1, atom2, length, K=None, group=1):
        """Give matching harmonic bonds a new equilibrium length (and constant) at their own
        time scale; the difference to the original potential goes to ``group``
        (systems.py:121-177)."""
        match = self._matcher(topology, residue, [atom1, atom2])
        changed = []
        for force in self.getForces():
            if isinstance(force, mm.HarmonicBondForce):
                for k in range(force.getNumBonds()):
                    i, j, r0, K0 = force.getBondParameters(k)
                    if match(i, j):
                        force.setBondParameters(k, i, j, length, K0 if K is None else K)
                        changed.append((i, j, r0, K0))
        if changed and self._special_bond_force is None:
            special = mm.CustomBondForce('0.5*(K0*(r - r0)^2 - Kn*(r - rn)^2)')
            for name in ('r0', 'K0', 'rn', 'Kn'):
                special.addPerBondParameter(name)
            special.setForceGroup(group)
            self.addForce(special)
            self._special_bond_force = special
        for i, j, r0, K0 in changed:
            self._special_bond_force.addBond(i, j, (r0, K0, length, K0 if K is None else K))

    def redefine_angle(self, topology, residue, atom1, atom2, atom3, angle, K=None, group=1):
        """Angle analogue of :meth:`redefine_bond` (systems.py:179-237)."""
        match = self._matcher(topology, residue, [atom1, atom2, atom3])
        changed = []
        for force in self.getForces():
            if isinstance(force, mm.HarmonicAngleForce):
                for n in range(force.getNumAngles()):
                    i, j, k, theta0, K0 = force.getAngleParameters(n)
                    if match(i, j, k):
                        force.setAngleParameters(n, i, j, k, angle, K0 if K is None else K)
                        changed.append((i, j, k, theta0, K0))
        if changed and self._special_angle_force is None:
            special = mm.CustomAngleForce('0.5*(K0*(theta - t0)^2 - Kn*(theta - tn)^2)')
            for name in ('t0', 'K0', 'tn', 'Kn'):
                special.addPerAngleParameter(name)
            special.setForceGroup(group)
            self.addForce(special)
            self._special_angle_force = special
        for i, j, k, theta0, K0 in changed:
            self._special_angle_force.addAngle(i, j, k, (theta0, K0, angle, K0 if K is None else K))


class AlchemicalSystem(mm.System):
    """A copy of ``system`` prepared for solvation free-energy / AFED runs (systems.py:318-410).

    The solute ``atoms`` are decoupled from the NonbondedForce (charge and epsilon zeroed, their
    mutual interactions kept as exceptions) and re-coupled to the rest of the system through a
    CustomNonbondedForce restricted to solute-solvent pairs by an interaction group, with the
    Beutler soft core ``4*lambda_vdw*epsilon*(1-x)/x^2; x=(r/sigma)^6+0.5*(1-lambda_vdw)`` and an
    energy derivative with respect to the context parameter ``lambda_vdw``.
    """

    def __init__(self, system, atoms, coupling='softcore', group=0, use_lrc=False):
        import itertools
        import math
        super().__init__()
        self._adopt(system)
        from .utils import findNonbondedForce, InputError
        if coupling != 'softcore':
            raise InputError('only the softcore coupling is available on the B200 engine')
        atoms = set(int(i) for i in atoms)
        nonbonded = self.getForce(findNonbondedForce(self))
        potential = 'U_softcore'
        potential += '; U_softcore = 4*lambda_vdw*epsilon*(1 - x)/x^2'
        potential += '; x = (r/sigma)^6 + 0.5*(1 - lambda_vdw)'
        potential += '; sigma = 0.5*(sigma1 + sigma2)'
        potential += '; epsilon = sqrt(epsilon1*epsilon2)'
        softcore = mm.CustomNonbondedForce(potential)
        if nonbonded.getNonbondedMethod() == mm.NonbondedForce.NoCutoff:
            softcore.setNonbondedMethod(mm.CustomNonbondedForce.NoCutoff)
        else:
            softcore.setNonbondedMethod(mm.CustomNonbondedForce.CutoffPeriodic)
        softcore.setCutoffDistance(nonbonded.getCutoffDistance())
        softcore.setUseSwitchingFunction(nonbonded.getUseSwitchingFunction())
        softcore.setSwitchingDistance(nonbonded.getSwitchingDistance())
        softcore.setUseLongRangeCorrection(use_lrc)
        softcore.addGlobalParameter('lambda_vdw', 1.0)
        softcore.addPerParticleParameter('sigma')
        softcore.addPerParticleParameter('epsilon')
        all_atoms = range(nonbonded.getNumParticles())
        for index in all_atoms:
            _, sigma, epsilon = nonbonded.getParticleParameters(index)
            softcore.addParticle([sigma, epsilon])
        for index in range(nonbonded.getNumExceptions()):
            i, j = nonbonded.getExceptionParameters(index)[:2]
            softcore.addExclusion(i, j)
        softcore.addInteractionGroup(atoms, set(all_atoms) - atoms)
        softcore.setForceGroup(group)
        softcore.addEnergyParameterDerivative('lambda_vdw')
        self.addForce(softcore)
        parameters = {}
        for i in atoms:
            parameters[i] = nonbonded.getParticleParameters(i)
            nonbonded.setParticleParameters(i, 0.0, 1.0, 0.0)
        exception_pairs = []
        for index in range(nonbonded.getNumExceptions()):
            i, j = nonbonded.getExceptionParameters(index)[:2]
            if {i, j} <= atoms:
                exception_pairs.append({i, j})
        for i, j in itertools.combinations(sorted(atoms), 2):
            if {i, j} not in exception_pairs:
                q1, sig1, eps1 = [unit.md_value(v) for v in parameters[i]]
                q2, sig2, eps2 = [unit.md_value(v) for v in parameters[j]]
                nonbonded.addException(i, j, q1*q2, (sig1 + sig2)/2, math.sqrt(eps1*eps2))
                softcore.addExclusion(i, j)  # keeps the exclusion list equal to the exception list


class ComputingSystem(_AtomsMM_System):
    """Virial contributions expressed as *energies* of an auxiliary system
    (systems.py:868-944):

    * group 0 (``_dispersion``): ``24 eps (2 (sig/r)^12 - (sig/r)^6)`` over pairs (same cutoff,
      switch and long-range-correction flags as the source NonbondedForce) and exceptions;
    * group 1 (``_bonded``): ``-K r (r - r0)`` per harmonic bond, ``-r dE/dr`` per custom bond;
    * group 2 (``_coulomb``): the source NonbondedForce with every epsilon zeroed, i.e. the
      Coulomb virial taken as the Coulomb energy.

    As in the reference, CustomNonbondedForce objects of the source system contribute nothing.
    """

    def __init__(self, system):
        super().__init__(system, copyForces=False)
        dispersion_group, bonded_group, coulomb_group = 0, 1, 2
        self._dispersion = 1 << dispersion_group
        self._bonded = 1 << bonded_group
        self._coulomb = 1 << coulomb_group
        lj_virial = '24*epsilon*(2*(sigma/r)^12-(sigma/r)^6)'
        for force in system.getForces():
            if isinstance(force, mm.NonbondedForce) and force.getNumParticles() > 0:
                source = copy.deepcopy(force)
                pairs = _forces._AtomsMM_CustomNonbondedForce(lj_virial)
                pairs.importFrom(source)
                pairs.setForceGroup(dispersion_group)
                self.addForce(pairs)
                exceptions = _forces._AtomsMM_CustomBondForce(lj_virial)
                exceptions.importFrom(source, extract=False)
                if exceptions.getNumBonds() > 0:
                    exceptions.setForceGroup(dispersion_group)
                    self.addForce(exceptions)
                for k in range(source.getNumParticles()):
                    charge = source.getParticleParameters(k)[0]
                    source.setParticleParameters(k, charge, 1.0, 0.0)
                for k in range(source.getNumExceptions()):
                    i, j, chargeprod = source.getExceptionParameters(k)[:3]
                    source.setExceptionParameters(k, i, j, chargeprod, 1.0, 0.0)
                source.setForceGroup(coulomb_group)
                source.setReciprocalSpaceForceGroup(coulomb_group)
                self.addForce(source)
            elif isinstance(force, mm.HarmonicBondForce) and force.getNumBonds() > 0:
                bonds = mm.CustomBondForce('-K*r*(r-r0)')
                bonds.addPerBondParameter('r0')
                bonds.addPerBondParameter('K')
                for k in range(force.getNumBonds()):
                    i, j, r0, K = force.getBondParameters(k)
                    bonds.addBond(i, j, [r0, K])
                bonds.setForceGroup(bonded_group)
                self.addForce(bonds)
            elif isinstance(force, mm.CustomBondForce) and force.getNumBonds() > 0:
                bonds = mm.CustomBondForce(self._virialExpression(force))
                for k in range(force.getNumPerBondParameters()):
                    bonds.addPerBondParameter(force.getPerBondParameterName(k))
                for k in range(force.getNumGlobalParameters()):
                    bonds.addGlobalParameter(force.getGlobalParameterName(k),
                                             force.getGlobalParameterDefaultValue(k))
                for k in range(force.getNumBonds()):
                    bonds.addBond(*force.getBondParameters(k))
                bonds.setUsesPeriodicBoundaryConditions(force.usesPeriodicBoundaryConditions())
                bonds.setForceGroup(bonded_group)
                self.addForce(bonds)

    def _virialExpression(self, force):
        """``-r dE/dr`` of a CustomBondForce energy, as a string (systems.py:934-944)."""
        energy = _expr.parse_inlined(force.getEnergyFunction())
        virial = ('neg', ('mul', ('var', 'r'), _expr.diff(energy, 'r')))
        return _expr.to_string(virial)

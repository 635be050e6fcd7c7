"""
Helper functions (same names and behaviour as the reference's ``atomsmm.utils``;
reference: src/atomsmm/utils.py:16-228).
"""

from collections import OrderedDict
from copy import deepcopy

import numpy as np

from . import mm
from . import unit

#: molar gas constant kB*N_A with simtk's CODATA-2006 values (utils.py:16)
kB = unit.BOLTZMANN_CONSTANT_kB*unit.AVOGADRO_CONSTANT_NA


class InputError(Exception):
    def __init__(self, msg):
        super(InputError, self).__init__('\033[1;31m' + msg + '\033[0m')


def countDegreesOfFreedom(system):
    """3 x (number of massive particles) - 3 - (number of constraints)  (utils.py:24-40)."""
    masses = getattr(system, '_masses', None)
    if masses is not None:          # this package's System: one pass over the stored floats
        massive = int(np.count_nonzero(np.asarray(masses, dtype=float) > 0))
    else:
        massive = sum(1 for i in range(system.getNumParticles())
                      if unit.md_value(system.getParticleMass(i)) > 0)
    return 3*massive - 3 - system.getNumConstraints()


def findNonbondedForce(system, position=0):
    """Index of the ``position``-th NonbondedForce of ``system`` (utils.py:43-62)."""
    found = [k for k, force in enumerate(system.getForces()) if isinstance(force, mm.NonbondedForce)]
    return found[position]


def hijackForce(system, index):
    """Remove force ``index`` from ``system`` and return a copy of it (utils.py:65-89)."""
    force = deepcopy(system.getForce(index))
    system.removeForce(index)
    return force


def globalParameters(force):
    return {force.getGlobalParameterName(k): force.getGlobalParameterDefaultValue(k)
            for k in range(force.getNumGlobalParameters())}


def _offset_parameters(force, count, getter):
    defaults = globalParameters(force)
    names = []
    for k in range(count):
        name = getter(k)[0]
        if name not in names:
            names.append(name)
    return OrderedDict((name, defaults[name]) for name in names)


def particleOffsetParameters(force):
    return _offset_parameters(force, force.getNumParticleParameterOffsets(), force.getParticleParameterOffset)


def exceptionOffsetParameters(force):
    return _offset_parameters(force, force.getNumExceptionParameterOffsets(), force.getExceptionParameterOffset)


def splitPotentialEnergy(system, topology, positions, **globals):
    """Potential energy split per Force object (utils.py:118-186).

    Every force gets its own group (a NonbondedForce two: ``Real-Space`` and
    ``Reciprocal-Space``); keys follow the reference: class name, then ``Name(1)``, ... for
    repeats, plus ``Total``.
    """
    from . import app
    clone = deepcopy(system)
    forces = clone.getForces()
    group = 0
    layout = []
    for force in forces:
        force.setForceGroup(group)
        entry = [force, group, None]
        group += 1
        if isinstance(force, mm.NonbondedForce):
            force.setReciprocalSpaceForceGroup(group)
            entry[2] = group
            group += 1
        layout.append(entry)
    simulation = app.Simulation(topology, clone, mm.VerletIntegrator(0.0), mm.Platform.getPlatformByName('B200'))
    simulation.context.setPositions(positions)
    for name, value in globals.items():
        simulation.context.setParameter(name, value)
    repeats = {}
    energy = OrderedDict()
    for force, direct, reciprocal in layout:
        label = force.__class__.__name__
        for base in force.__class__.__mro__:
            if base.__module__ == mm.__name__ and base is not mm.Force:
                label = base.__name__
                break
        if label == 'NonbondedForce':
            label = 'Real-Space'
        value = simulation.context.getState(getEnergy=True, groups={direct}).getPotentialEnergy()
        first = label not in repeats
        repeats[label] = 0 if first else repeats[label] + 1
        energy[label if first else '%s(%d)' % (label, repeats[label])] = value
        if reciprocal is not None:
            value = simulation.context.getState(getEnergy=True, groups={reciprocal}).getPotentialEnergy()
            energy['Reciprocal-Space' if first else 'Reciprocal-Space(%d)' % repeats[label]] = value
    energy['Total'] = sum(energy.values(), 0.0*unit.kilojoules_per_mole)
    return energy


def evaluateForce(force, positions, boxVectors=None):
    """Energy of a single Force for given coordinates (utils.py:189-228)."""
    system = mm.System()
    for _ in range(len(positions)):
        system.addParticle(0)
    if boxVectors is not None:
        system.setDefaultPeriodicBoxVectors(*boxVectors)
    system.addForce(deepcopy(force))
    context = mm.Context(system, mm.CustomIntegrator(0), mm.Platform.getPlatformByName('B200'))
    context.setPositions(positions)
    return context.getState(getEnergy=True).getPotentialEnergy()

"""System builders shared by CPU and GPU tests (the constructions of the reference's tests)."""

import numpy as np

import atomsmm_b200 as atomsmm
from atomsmm_b200 import app, mm, unit

import fixtures

A = unit.angstroms


def positions_of(structure):
    return structure.getPositions(asNumpy=True).value_in_unit(unit.nanometer)


def water_near(adjustment, rc=10*A, rs=9.5*A):
    """tests/test_respa_forces.py:11-26"""
    pdb, ff = fixtures.load('q-SPC-FW')
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffPeriodic)
    force = atomsmm.NearNonbondedForce(rc, rs, adjustment)
    force.importFrom(atomsmm.hijackForce(system, atomsmm.findNonbondedForce(system))).addTo(system)
    return system, pdb, force


def water_damped(degree):
    """tests/test_DampedSmoothedForce.py:11-27"""
    pdb, ff = fixtures.load('q-SPC-FW')
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffPeriodic)
    force = atomsmm.DampedSmoothedForce(0.29/A, 10*A, 9.5*A, degree=degree)
    force.importFrom(atomsmm.hijackForce(system, atomsmm.findNonbondedForce(system))).addTo(system)
    return system, pdb, force


def il_exceptions():
    """tests/test_ExceptionNonbondedForce.py:11-24"""
    pdb, ff = fixtures.load('emim_BCN4_Jiung2014')
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffPeriodic)
    force = atomsmm.forces.NonbondedExceptionsForce()
    force.importFrom(atomsmm.hijackForce(system, atomsmm.findNonbondedForce(system))).addTo(system)
    return system, pdb, force


def flexible(case, method=app.PME, **kwargs):
    """tests/test_computers.py:11-19"""
    pdb, ff = fixtures.load(case)
    system = ff.createSystem(pdb.topology, nonbondedMethod=method, constraints=None, rigidWater=False,
                             removeCMMotion=False, **kwargs)
    return system, pdb


def respa_water(method=app.CutoffPeriodic, rcut_in=7*A, rswitch_in=5*A, **kwargs):
    """RESPASystem on flexible q-SPC-FW (BASELINE config 1; tests/test_systems.py:131-135)"""
    system, pdb = flexible('q-SPC-FW', method)
    return atomsmm.RESPASystem(system, rcut_in, rswitch_in, **kwargs), pdb


def replicate(system, positions, box, reps, jitter=0.002, seed=1):
    """Tile a periodic system reps^3 times (BASELINE configs 2, 3, 5): new System with the same
    forces on shifted copies, every atom displaced by uniform(-jitter, jitter) nm."""
    import copy
    n = system.getNumParticles()
    copies = reps**3
    big = mm.System()
    for _ in range(copies):
        for i in range(n):
            big.addParticle(system.getParticleMass(i))
    b = np.asarray(box, dtype=float)
    big.setDefaultPeriodicBoxVectors(mm.Vec3(b[0]*reps, 0, 0), mm.Vec3(0, b[1]*reps, 0), mm.Vec3(0, 0, b[2]*reps))
    shifts = [(ix, iy, iz) for ix in range(reps) for iy in range(reps) for iz in range(reps)]
    pos = np.concatenate([positions + np.array(s)*b for s in shifts], axis=0)
    rng = np.random.default_rng(seed)
    pos = pos + rng.uniform(-jitter, jitter, size=pos.shape)
    for force in system.getForces():
        new = copy.deepcopy(force)
        if isinstance(force, mm.NonbondedForce):
            new._particles = [list(p) for _ in range(copies) for p in force._particles]
            new._exceptions = [[e[0] + c*n, e[1] + c*n] + list(e[2:]) for c in range(copies) for e in force._exceptions]
            new._exception_index = {(min(e[0], e[1]), max(e[0], e[1])): k for k, e in enumerate(new._exceptions)}
        elif isinstance(force, mm.CustomNonbondedForce):
            new._particles = [list(p) for _ in range(copies) for p in force._particles]
            new._exclusions = [(i + c*n, j + c*n) for c in range(copies) for i, j in force._exclusions]
        elif isinstance(force, (mm.HarmonicBondForce, mm.CustomBondForce)):
            new._bonds = [[bd[0] + c*n, bd[1] + c*n] + list(bd[2:]) for c in range(copies) for bd in force._bonds]
        elif isinstance(force, (mm.HarmonicAngleForce, mm.CustomAngleForce)):
            new._angles = [[a[0] + c*n, a[1] + c*n, a[2] + c*n] + list(a[3:]) for c in range(copies) for a in force._angles]
        elif isinstance(force, mm.PeriodicTorsionForce):
            new._torsions = [[t[0] + c*n, t[1] + c*n, t[2] + c*n, t[3] + c*n] + list(t[4:]) for c in range(copies)
                             for t in force._torsions]
        big.addForce(new)
    return big, pos


def subset(case, nresidues):
    """(structure, forcefield) with only the first ``nresidues`` residues of a data set, in the original
    box: a small cluster for CPU-only tests of host logic (not a physical liquid)."""
    structure, ff = fixtures.load(case)
    record = fixtures._CACHE[case]
    natoms = sum(len(r[2]) for r in record['residues'][:nresidues])
    small = dict(record, residues=record['residues'][:nresidues], positions=record['positions'][:natoms])
    return fixtures.Structure(small), ff

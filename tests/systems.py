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
    forces on shifted copies, every atom displaced by uniform(-jitter, jitter) nm.  Tables are tiled
    with numpy and kept columnar (mm.PackedRows): config 5 has 4.2 M atoms."""
    import copy
    n = system.getNumParticles()
    copies = reps**3
    big = mm.System()
    big._masses = np.tile(np.asarray(system._masses, dtype=np.float64), copies).tolist()
    b = np.asarray(box, dtype=float)
    big.setDefaultPeriodicBoxVectors(mm.Vec3(b[0]*reps, 0, 0), mm.Vec3(0, b[1]*reps, 0), mm.Vec3(0, 0, b[2]*reps))
    shifts = np.array([(ix, iy, iz) for ix in range(reps) for iy in range(reps) for iz in range(reps)], dtype=float)
    pos = (np.asarray(positions)[None, :, :] + (shifts*b)[:, None, :]).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    pos = pos + rng.uniform(-jitter, jitter, size=pos.shape)
    offsets = (np.arange(copies, dtype=np.int64)*n)[:, None, None]

    def tile(rows, nindex, nested=False):
        if len(rows) == 0:
            return []
        idx = mm.index_columns(rows, nindex).astype(np.int64)
        val = mm.value_columns(rows, nindex)
        idx = (idx[None, :, :] + offsets).reshape(-1, nindex) if nindex else None
        return mm.PackedRows(idx, np.tile(val, (copies, 1)), nested=nested)

    for force in system.getForces():
        tables = {}
        if isinstance(force, mm.NonbondedForce):
            tables = dict(_particles=(0, False), _exceptions=(2, False))
        elif isinstance(force, mm.CustomNonbondedForce):
            tables = dict(_particles=(0, False), _exclusions=(2, False))
        elif isinstance(force, mm.HarmonicBondForce):
            tables = dict(_bonds=(2, False))
        elif isinstance(force, mm.CustomBondForce):
            tables = dict(_bonds=(2, True))
        elif isinstance(force, mm.HarmonicAngleForce):
            tables = dict(_angles=(3, False))
        elif isinstance(force, mm.CustomAngleForce):
            tables = dict(_angles=(3, True))
        elif isinstance(force, mm.PeriodicTorsionForce):
            tables = dict(_torsions=(4, False))
        saved = {name: force.__dict__[name] for name in tables}
        for name in tables:
            force.__dict__[name] = []          # do not deep-copy what is about to be replaced
        try:
            new = copy.deepcopy(force)
        finally:
            force.__dict__.update(saved)
        for name, (nindex, nested) in tables.items():
            new.__dict__[name] = tile(saved[name], nindex, nested)
        if isinstance(force, mm.NonbondedForce):
            new._exception_index = {}
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

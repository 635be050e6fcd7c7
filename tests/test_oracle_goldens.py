"""
Pins the ORACLE (oracle/refmath.py) to the reference's own known-answer tests: every single-point
golden of /root/reference/tests that does not depend on OpenMM's random stream.
Tolerance is the reference's own pytest.approx default (rel 1e-6) unless stated.
"""

import numpy as np
import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import app, mm, unit
from oracle import refmath

import systems
from systems import positions_of

REL = 1e-6


@pytest.mark.parametrize('adjustment,golden', [(None, -24955.845391462222), ('shift', -26451.885982885935),
                                               ('force-switch', -26516.68871844118)])
def test_near_force(adjustment, golden):
    # reference: tests/test_respa_forces.py:29-38
    system, pdb, _ = systems.water_near(adjustment)
    assert refmath.evaluate_system(system, positions_of(pdb)).energy == pytest.approx(golden, rel=REL)


@pytest.mark.parametrize('degree,golden', [(1, -25074.251664020387), (2, -25074.342992954276)])
def test_damped_smoothed(degree, golden):
    # reference: tests/test_DampedSmoothedForce.py:30-35
    system, pdb, _ = systems.water_damped(degree)
    assert refmath.evaluate_system(system, positions_of(pdb)).energy == pytest.approx(golden, rel=REL)


def test_exceptions():
    # reference: tests/test_ExceptionNonbondedForce.py:11-24
    system, pdb, _ = systems.il_exceptions()
    assert refmath.evaluate_system(system, positions_of(pdb)).energy == pytest.approx(-27616.298459208883, rel=REL)


def _molecules(system):
    from atomsmm_b200.engine import _molecules
    return _molecules(system)[0]


@pytest.mark.parametrize('case,goldens', [
    ('q-SPC-FW', (-11661.677650154408, -58.64837784125407, -5418.629781093525, -554.9525554206972)),
    ('emim_BCN4_Jiung2014', (-22827.477810819175, -282.7243180164338, -23272.958585794207, -3283.563262288828))])
def test_pressure_computer_goldens(case, goldens):
    # reference: tests/test_computers.py:22-37, 59-74 (PME, flexible, bath temperature 300 K)
    system, pdb = systems.flexible(case)
    pos = positions_of(pdb)
    computing = atomsmm.ComputingSystem(system)
    W = sum(refmath.evaluate_system(computing, pos, groups={g}).energy for g in (0, 1, 2))
    assert W == pytest.approx(goldens[0], rel=REL)
    box = refmath.system_box(system)
    kT = 1.3806504e-23*6.02214179e23*300/1000.0
    n = system.getNumParticles()
    to_atm = 1e3/6.02214179e23/1e-27/101325.0
    # the atomic pressure is a small difference of large terms (3NkT = +11493 vs W = -11662 for
    # water): the 2e-8 relative error of W, which comes from OpenMM's own numerical long-range
    # integration, is amplified to 1.5e-6 here
    assert (3*n*kT + W)/(3*np.prod(box))*to_atm == pytest.approx(goldens[1], rel=3e-6)
    forces = refmath.evaluate_system(system, pos).forces
    mol = _molecules(system)
    mass = np.array([system.getParticleMass(i).value_in_md_units() for i in range(n)])
    mol_mass = np.bincount(mol, mass)
    rcm = np.stack([np.bincount(mol, mass*pos[:, k])/mol_mass for k in range(3)], 1)
    fcm = np.stack([np.bincount(mol, forces[:, k]) for k in range(3)], 1)
    Wm = W + np.sum(rcm*fcm) - np.sum(pos*forces)
    assert Wm == pytest.approx(goldens[2], rel=REL)
    assert (3*len(mol_mass)*kT + Wm)/(3*np.prod(box))*to_atm == pytest.approx(goldens[3], rel=REL)


def test_respa_system_goldens():
    # reference: tests/test_systems.py:131-152
    pdb_system, pdb = systems.flexible('q-SPC-FW', app.PME)
    nb = pdb_system.getForce(atomsmm.findNonbondedForce(pdb_system))
    nb.setUseSwitchingFunction(True)
    nb.setSwitchingDistance(9*unit.angstroms)          # readSystem(), test_systems.py:20-22 (rswitch 9 A)
    respa = atomsmm.RESPASystem(pdb_system, 7*unit.angstroms, 5*unit.angstroms)
    respa.redefine_bond(pdb.topology, 'HOH', 'H[1-2]', 'O', 1.05*unit.angstroms)
    respa.redefine_angle(pdb.topology, 'HOH', 'H[1-2]', 'O', 'H[1-2]', 113*unit.degrees)
    pos = positions_of(pdb)
    goldens = {'HarmonicBondForce': 3665.684696323676, 'HarmonicAngleForce': 1811.197218501007,
               'PeriodicTorsionForce': 0.0, 'Real-Space': 84694.39953220935,
               'Reciprocal-Space': -111582.71281220087, 'CustomNonbondedForce': -25531.129587235544,
               'CustomNonbondedForce(1)': 25531.129587235544, 'CustomBondForce': 0.0,
               'CustomBondForce(1)': -1175.253817235862, 'CustomAngleForce': -305.0221912655623}
    box = refmath.system_box(respa)
    got, repeats = {}, {}
    for force in respa.getForces():
        label = refmath._kind(force)
        if label == 'NonbondedForce':
            got['Real-Space'] = refmath.eval_nonbonded(force, pos, box, 'direct').energy
            got['Reciprocal-Space'] = refmath.eval_nonbonded(force, pos, box, 'reciprocal').energy
            continue
        evaluator = {'CustomNonbondedForce': refmath.eval_custom_nonbonded, 'CustomBondForce': refmath.eval_custom_bond,
                     'CustomAngleForce': refmath.eval_custom_angle, 'HarmonicBondForce': refmath.eval_harmonic_bond,
                     'HarmonicAngleForce': refmath.eval_harmonic_angle,
                     'PeriodicTorsionForce': refmath.eval_periodic_torsion}[label]
        first = label not in repeats
        repeats[label] = 0 if first else repeats[label] + 1
        got[label if first else '%s(%d)' % (label, repeats[label])] = evaluator(force, pos, box).energy
    for key, value in goldens.items():
        assert got[key] == pytest.approx(value, rel=REL, abs=1e-9), key
    assert sum(got.values()) == pytest.approx(-22891.707373668243, rel=REL)


def test_forces_are_gradients():
    """Oracle self-consistency: analytic forces equal -dE/dx by central differences."""
    system, pdb = systems.flexible('emim_BCN4_Jiung2014', app.CutoffPeriodic)
    pos = positions_of(pdb)
    base = refmath.evaluate_system(system, pos)
    rng = np.random.default_rng(0)
    for atom in rng.choice(len(pos), 4, replace=False):
        for k in range(3):
            h = 1e-6
            p = pos.copy(); p[atom, k] += h
            m = pos.copy(); m[atom, k] -= h
            numeric = -(refmath.evaluate_system(system, p).energy - refmath.evaluate_system(system, m).energy)/(2*h)
            assert base.forces[atom, k] == pytest.approx(numeric, rel=2e-5, abs=2e-3)

"""
GPU parity tests (run with -m gpu on the B200 box): the CUDA engine, called through the public
atomsmm API -> ctypes C ABI, against (a) the reference's golden energies and (b) the float64
oracle on the same inputs.  Tolerances (BASELINE.json north_star): energies and virials 1e-6
relative, per-atom forces 1e-5 relative RMS, interacting pair sets bit-exact.
"""

import numpy as np
import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import app, mm, unit

import systems
from systems import positions_of

pytestmark = pytest.mark.gpu

E_REL = 1e-6
F_RMS = 1e-5


def rel_rms(a, b):
    return float(np.sqrt(np.sum((a - b)**2)/np.sum(b**2)))


def mix64(z):
    z = (z + np.uint64(0x9e3779b97f4a7c15)).astype(np.uint64)
    z = ((z ^ (z >> np.uint64(30)))*np.uint64(0xbf58476d1ce4e5b9)).astype(np.uint64)
    z = ((z ^ (z >> np.uint64(27)))*np.uint64(0x94d049bb133111eb)).astype(np.uint64)
    return z ^ (z >> np.uint64(31))


def pair_checksum(i, j):
    lo = np.minimum(i, j).astype(np.uint64)
    hi = np.maximum(i, j).astype(np.uint64)
    with np.errstate(over='ignore'):
        return int(np.sum(mix64((lo << np.uint64(32)) | hi), dtype=np.uint64))


def make_context(system, pdb, platform, **properties):
    context = mm.Context(system, mm.VerletIntegrator(0.0), platform, properties)
    context.setPositions(pdb.positions)
    return context


def check_against_oracle(system, pdb, platform, groups=None, e_rel=E_REL, f_rms=F_RMS):
    from oracle import refmath
    context = make_context(system, pdb, platform)
    pos = positions_of(pdb)
    g = -1 if groups is None else groups
    state = context.getState(getEnergy=True, getForces=True, groups=g)
    ref = refmath.evaluate_system(system, pos, groups=groups)
    energy = state.getPotentialEnergy().value_in_unit(unit.kilojoules_per_mole)
    forces = state.getForces(asNumpy=True).value_in_unit(unit.kilojoules_per_mole/unit.nanometer)
    assert energy == pytest.approx(ref.energy, rel=e_rel, abs=1e-6)
    if np.sum(ref.forces**2) > 0:
        assert rel_rms(forces, ref.forces) < f_rms
    return context, state, ref


@pytest.mark.parametrize('adjustment,golden', [(None, -24955.845391462222), ('shift', -26451.885982885935),
                                               ('force-switch', -26516.68871844118)])
def test_near_force(cuda_platform, adjustment, golden):
    # reference: tests/test_respa_forces.py:11-38
    system, pdb, force = systems.water_near(adjustment)
    context, state, ref = check_against_oracle(system, pdb, cuda_platform)
    assert state.getPotentialEnergy().value_in_unit(unit.kilojoules_per_mole) == pytest.approx(golden, rel=E_REL)


@pytest.mark.parametrize('degree,golden', [(1, -25074.251664020387), (2, -25074.342992954276)])
def test_damped_smoothed(cuda_platform, degree, golden):
    # reference: tests/test_DampedSmoothedForce.py:11-35
    system, pdb, force = systems.water_damped(degree)
    context, state, ref = check_against_oracle(system, pdb, cuda_platform)
    assert state.getPotentialEnergy().value_in_unit(unit.kilojoules_per_mole) == pytest.approx(golden, rel=E_REL)


def test_exceptions(cuda_platform):
    # reference: tests/test_ExceptionNonbondedForce.py:11-24
    system, pdb, force = systems.il_exceptions()
    context, state, ref = check_against_oracle(system, pdb, cuda_platform)
    assert state.getPotentialEnergy().value_in_unit(unit.kilojoules_per_mole) == pytest.approx(-27616.298459208883, rel=E_REL)


@pytest.mark.parametrize('adjustment', [None, 'shift', 'force-switch'])
def test_pair_set_bit_exact(cuda_platform, adjustment):
    """Interacting pair set {(i,j): r < rc0 in float64 minimum image, not excluded} equals the
    oracle's, pair for pair."""
    from oracle import refmath
    system, pdb, force = systems.water_near(adjustment, 7*systems.A, 5*systems.A)
    context = make_context(system, pdb, cuda_platform)
    ref = refmath.eval_custom_nonbonded(force, positions_of(pdb), refmath.system_box(system), want_pairs=True)
    i, j, r = ref.pairs
    count, checksum, pairs = context.pair_set(force, want_pairs=True)
    assert count == len(i)
    assert checksum == pair_checksum(i, j)
    ours = set(map(tuple, np.sort(pairs, axis=1).tolist()))
    theirs = set(zip(np.minimum(i, j).tolist(), np.maximum(i, j).tolist()))
    assert ours == theirs


@pytest.mark.parametrize('case', ['q-SPC-FW', 'emim_BCN4_Jiung2014'])
def test_respa_system_groups(cuda_platform, case):
    """RESPASystem over a reaction-field NonbondedForce: every force group against the oracle
    (group 0 bonded + 1-4, group 1 near force-switch, group 2 full LJ+RF, group 31 -near)."""
    system, pdb = systems.flexible(case, app.CutoffPeriodic)
    respa = atomsmm.RESPASystem(system, 7*systems.A, 5*systems.A)
    for groups in ({0}, {1}, {2}, {31}, None):
        check_against_oracle(respa, pdb, cuda_platform, groups=groups)


def test_far_force_identity(cuda_platform):
    """FarNonbondedForce + NearNonbondedForce == full NonbondedForce (reaction field), the identity
    of tests/test_respa_forces.py:41-79 with CutoffPeriodic instead of PME."""
    pdb, ff = systems.fixtures.load('q-SPC-FW')
    for adjustment in (None, 'shift', 'force-switch'):
        system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffPeriodic)
        nb = atomsmm.hijackForce(system, atomsmm.findNonbondedForce(system))
        inner = atomsmm.NearNonbondedForce(7*systems.A, 6.5*systems.A, adjustment)
        inner.importFrom(nb).addTo(system)
        outer = atomsmm.FarNonbondedForce(inner, 10*systems.A, 9.5*systems.A).setForceGroup(2)
        outer.importFrom(nb).addTo(system)
        total = atomsmm.splitPotentialEnergy(system, pdb.topology, pdb.positions)['Total']
        refsys = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffPeriodic, nonbondedCutoff=10*systems.A)
        force = refsys.getForce(refsys.getNumForces() - 2)
        force.setUseSwitchingFunction(True)
        force.setSwitchingDistance(9.5*systems.A)
        reference = atomsmm.splitPotentialEnergy(refsys, pdb.topology, pdb.positions)['Total']
        assert total/total.unit == pytest.approx(reference/reference.unit, rel=E_REL)


def test_replicated_box_against_oracle(cuda_platform):
    """2x2x2 replica of the water box (12 288 atoms, jittered): staged-image path of the pair
    kernel and multi-cell lists, against the oracle's KD-tree pair search."""
    from oracle import refmath
    system, pdb = systems.flexible('q-SPC-FW', app.CutoffPeriodic)
    respa = atomsmm.RESPASystem(system, 7*systems.A, 5*systems.A)
    big, pos = systems.replicate(respa, positions_of(pdb), refmath.system_box(system), 2)
    context = mm.Context(big, mm.VerletIntegrator(0.0), cuda_platform)
    context.setPositions(pos)
    for groups in ({1}, {2}):
        state = context.getState(getEnergy=True, getForces=True, groups=groups)
        ref = refmath.evaluate_system(big, pos, groups=groups)
        assert state.getPotentialEnergy().value_in_unit(unit.kilojoules_per_mole) == pytest.approx(ref.energy, rel=E_REL)
        forces = state.getForces(asNumpy=True).value_in_unit(unit.kilojoules_per_mole/unit.nanometer)
        assert rel_rms(forces, ref.forces) < F_RMS

"""
The engine's spatial order (SURVEY 2.1 kernel K1) is computed on the device (csrc/order.cu: one Hilbert key per
molecule, stable radix sort, gathers of every static per-atom table).  It must be EXACTLY the order the pure host
function b2_hilbert_index defines -- whole molecules contiguous, molecules sorted by (key of the first atom, molecule
id) -- both at the first setPositions and after a re-ordering in the middle of a run, and the re-ordered state must
be the same physical state (positions, velocities and forces in the caller's numbering are unchanged).
"""

import numpy as np
import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import engine, mm, unit

import systems

pytestmark = pytest.mark.gpu


def _expected_order(pos, box, molecule):
    molecule = np.asarray(molecule)
    first = {}
    for i, m in enumerate(molecule):
        first.setdefault(int(m), i)
    keys = sorted((engine.hilbert_index(pos[a], box), m) for m, a in first.items())
    members = {}
    for i, m in enumerate(molecule):
        members.setdefault(int(m), []).append(i)
    return np.array([i for _, m in keys for i in members[m]], dtype=np.int32)


def test_device_order_is_the_hilbert_order_of_the_host_function(cuda_platform):
    respa, pdb = systems.respa_water()
    pos = systems.positions_of(pdb)
    integrator = atomsmm.RespaPropagator([4, 2, 1]).integrator(4*unit.femtoseconds)
    context = mm.Context(respa, integrator, cuda_platform)
    context.setPositions(pos)
    context.setVelocitiesToTemperature(300*unit.kelvin, 7)
    box = np.asarray(context._box, dtype=np.float64)
    molecule = context._molecule
    order = context.spatial_order()
    assert sorted(order.tolist()) == list(range(len(pos)))
    assert np.array_equal(order, _expected_order(np.asarray(pos, dtype=np.float64), box, molecule))

    # a re-ordering in the middle of a run: shift every molecule by a lattice-incommensurate vector (the keys
    # change), keep the velocities; the state seen through the API must be that configuration, in the new order
    state = context.getState(getPositions=True, getVelocities=True)
    moved = state._positions + np.array([0.7, 1.3, 0.4])
    context.setPositions(moved)
    again = context.getState(getPositions=True, getVelocities=True, getForces=True)
    assert np.array_equal(again._positions, moved)
    assert np.array_equal(again._velocities, state._velocities)
    order2 = context.spatial_order()
    assert np.array_equal(order2, _expected_order(moved, box, molecule))
    assert not np.array_equal(order, order2)
    # forces are those of a fresh context at the same configuration
    fresh = mm.Context(respa, atomsmm.RespaPropagator([4, 2, 1]).integrator(4*unit.femtoseconds), cuda_platform)
    fresh.setPositions(moved)
    reference = fresh.getState(getForces=True)
    assert np.max(np.abs(again._forces - reference._forces)) <= 1e-3*np.max(np.abs(reference._forces))

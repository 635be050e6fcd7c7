"""The C/OpenMP oracle (oracle/c/oracle.c) against the pinned numpy/sympy oracle, and its RESPA
driver against the generic step-program interpreter."""

import numpy as np
import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import app, unit
from oracle import cport, interp, refmath

import systems
from systems import positions_of


@pytest.mark.parametrize('case', ['q-SPC-FW', 'emim_BCN4_Jiung2014'])
def test_c_port_matches_numpy_oracle(case):
    system, pdb = systems.flexible(case, app.CutoffPeriodic)
    respa = atomsmm.RESPASystem(system, 7*systems.A, 5*systems.A)
    pos = positions_of(pdb)
    port = cport.CPort(respa)
    for groups in ({0}, {1}, {2}, {31}):
        forces, energy, virial = port.evaluate(pos, groups)
        ref = refmath.evaluate_system(respa, pos, groups=groups)
        assert energy == pytest.approx(ref.energy, rel=1e-10, abs=1e-8)
        assert np.max(np.abs(forces - ref.forces)) < 1e-7*max(1.0, np.max(np.abs(ref.forces)))


@pytest.mark.parametrize('degree', [1, 2])
def test_c_port_damped(degree):
    system, pdb, force = systems.water_damped(degree)
    pos = positions_of(pdb)
    forces, energy, virial = cport.CPort(system).evaluate(pos)
    ref = refmath.evaluate_system(system, pos)
    assert energy == pytest.approx(ref.energy, rel=1e-10)
    assert virial == pytest.approx(ref.virial, rel=1e-9)
    assert np.max(np.abs(forces - ref.forces)) < 1e-7*np.max(np.abs(ref.forces))


def test_c_respa_matches_interpreter():
    """orc_respa hand-codes RespaPropagator([4,2,1]) + SuzukiYoshida(NoseHoover(nloops=2), 3): it must
    follow the generic interpreter running the program atomsmm emits."""
    respa, pdb = systems.respa_water()
    pos = positions_of(pdb)
    rng = np.random.default_rng(5)
    mass = np.array([respa.getParticleMass(i).value_in_md_units() for i in range(len(pos))])
    vel = rng.standard_normal(pos.shape)*np.sqrt(2.494/mass)[:, None]
    dof = atomsmm.countDegreesOfFreedom(respa)
    K, fs = unit.kelvin, unit.femtoseconds
    nh = atomsmm.NoseHooverPropagator(300*K, dof, 100*fs, 2)
    integrator = atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([4, 2, 1]),
                                                 atomsmm.SuzukiYoshidaPropagator(nh, 3)).integrator(4*fs)
    reference = interp.Interpreter(respa, integrator, pos, vel)
    reference.step(2)
    port = cport.CPort(respa)
    x, v, p_eta = port.respa(pos, vel, 2, 0.004, 4, 2, (2, reference.globals['LkT'], reference.globals['Q'], 0.0))
    assert np.max(np.abs(x - reference.x)) < 1e-9
    assert np.max(np.abs(v - reference.v)) < 1e-7
    assert p_eta == pytest.approx(reference.globals['p_eta'], rel=1e-9)

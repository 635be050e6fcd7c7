"""
``PressureComputer``: virial and pressure of a configuration, evaluated on the CUDA engine.

Same public methods as the reference class (reference: src/atomsmm/computers.py:46-246; rows
a9/a10 of SURVEY 8a).  Two definitions of the virial are available:

``pressure_mode='reference'`` (default)
    exactly the reference's: a second context over ``ComputingSystem`` whose three force-group
    *energies* are the dispersion, bonded and Coulomb virials (computers.py:145-172).  This is
    what the reference's goldens pin.  As in the reference, CustomNonbondedForce objects of the
    simulated system contribute nothing and the Coulomb virial is the Coulomb energy.
``pressure_mode='pair'``
    the true pair virial  W = sum_pairs r_ij . F_ij  of every two-body term of the simulated
    system, accumulated by the engine's float64 energy kernels in the same pass as the energies.
    This is the physically meaningful number for systems built with DampedSmoothedForce or
    Near/Far forces, for which the reference reports only the bonded part.
"""

import itertools

import numpy as np

from . import engine
from . import mm
from . import unit
from .systems import ComputingSystem
from .unit import md_value as _md


class _MoleculeTotalizer(object):
    """Per-molecule sums via index arrays (the reference uses scipy CSR matrices,
    computers.py:22-43)."""

    def __init__(self, context, topology):
        molecules = context.getMolecules()
        self.nmols = len(molecules)
        self.natoms = sum(len(m) for m in molecules)
        self.index = np.empty(self.natoms, dtype=np.int64)
        for k, molecule in enumerate(molecules):
            self.index[list(molecule)] = k
        system = context.getSystem()
        self.mass = np.array([_md(system.getParticleMass(i)) for i in range(self.natoms)])
        self.molMass = np.bincount(self.index, self.mass, self.nmols)
        self.massFrac = self.mass/self.molMass[self.index]
        residues = {}
        if topology is not None:
            for atom in topology.atoms():
                residues[int(atom.index)] = atom.residue.name
        self.residues = [residues.get(m[0], '') for m in molecules]

    def sum(self, values):
        """[natoms, 3] -> [nmols, 3] plain sums."""
        return np.stack([np.bincount(self.index, values[:, k], self.nmols) for k in range(3)], axis=1)

    def weighted(self, values):
        """[natoms, 3] -> [nmols, 3] mass-weighted means."""
        return self.sum(values*self.massFrac[:, None])


class PressureComputer(engine.Context):
    """
    Parameters
    ----------
        system : System
        topology : app.Topology
        platform : Platform
        properties : dict, optional
        temperature : unit.Quantity, optional
            If given, kinetic terms are replaced by their equipartition values.
        pressure_mode : 'reference' | 'pair'
    """

    def __init__(self, system, topology, platform, properties=dict(), temperature=None, pressure_mode='reference'):
        if pressure_mode not in ('reference', 'pair'):
            raise ValueError('pressure_mode must be "reference" or "pair"')
        self._pressure_mode = pressure_mode
        self._computing_system = ComputingSystem(system) if pressure_mode == 'reference' else system
        super().__init__(self._computing_system, mm.CustomIntegrator(0), platform, properties)
        self._mols = _MoleculeTotalizer(self, topology)
        self._kT = None if temperature is None else unit.MOLAR_GAS_CONSTANT_R*temperature
        self._make_obsolete()

    def _get_forces(self, groups):
        return self.getState(getForces=True, groups=groups).getForces(asNumpy=True)

    def _get_positions(self):
        return self.getState(getPositions=True).getPositions(asNumpy=True)

    def _get_potential(self, groups):
        return self.getState(getEnergy=True, groups=groups).getPotentialEnergy()

    def _get_velocities(self):
        return self.getState(getVelocities=True).getVelocities(asNumpy=True)

    def _get_volume(self):
        box = self.getState().getPeriodicBoxVectors()
        return box[0][0]*box[1][1]*box[2][2]*unit.AVOGADRO_CONSTANT_NA

    def _make_obsolete(self):
        self._bond_virial = None
        self._coulomb_virial = None
        self._dispersion_virial = None
        self._pair_virial = None
        self._molecular_kinetic_energy = None

    def get_atomic_pressure(self):
        """P = (2K + W)/(3V) with the unconstrained atomic virial (computers.py:97-123)."""
        if self._kT is None:
            velocities = self._get_velocities().value_in_unit(unit.nanometers/unit.picosecond)
            dNkT = float(np.sum(self._mols.mass*np.sum(velocities**2, axis=1)))*unit.kilojoules_per_mole
        else:
            dNkT = 3*self._mols.natoms*self._kT
        pressure = (dNkT + self.get_atomic_virial())/(3*self._get_volume())
        return pressure.in_units_of(unit.atmospheres)

    def get_atomic_virial(self):
        """W = -sum_ij r_ij E'(r_ij) (computers.py:125-145)."""
        if self._pressure_mode == 'pair':
            if self._pair_virial is None:
                state = self.getState(getEnergy=True)
                self._pair_virial = state._virial*unit.kilojoules_per_mole
            return self._pair_virial
        return self.get_bond_virial() + self.get_coulomb_virial() + self.get_dispersion_virial()

    def _component(self, attribute, mask_name):
        if self._pressure_mode != 'reference':
            raise mm.OpenMMException('virial components are only defined in pressure_mode="reference"')
        if getattr(self, attribute) is None:
            setattr(self, attribute, self._get_potential(getattr(self._computing_system, mask_name)))
        return getattr(self, attribute)

    def get_bond_virial(self):
        return self._component('_bond_virial', '_bonded')

    def get_coulomb_virial(self):
        return self._component('_coulomb_virial', '_coulomb')

    def get_dispersion_virial(self):
        return self._component('_dispersion_virial', '_dispersion')

    def get_molecular_kinetic_energy(self):
        if self._molecular_kinetic_energy is None:
            velocities = self._get_velocities().value_in_unit(unit.nanometers/unit.picosecond)
            vcm = self._mols.weighted(velocities)
            self._molecular_kinetic_energy = 0.5*float(np.sum(self._mols.molMass*np.sum(vcm**2, axis=1))) * \
                unit.kilojoules_per_mole
        return self._molecular_kinetic_energy

    def get_molecular_pressure(self, forces):
        """P = (2 K_mol + W_mol)/(3V) (computers.py:182-209)."""
        if self._kT is None:
            dNkT = 2.0*self.get_molecular_kinetic_energy()
        else:
            dNkT = 3*self._mols.nmols*self._kT
        pressure = (dNkT + self.get_molecular_virial(forces))/(3*self._get_volume())
        return pressure.in_units_of(unit.atmospheres)

    def get_molecular_virial(self, forces):
        """W_mol = W - sum_i (r_i - r_i^cm) . F_i (computers.py:211-240)."""
        f = np.asarray(forces.value_in_unit(unit.kilojoules_per_mole/unit.nanometers), dtype=np.float64)
        r = self._get_positions().value_in_unit(unit.nanometers)
        fcm = self._mols.sum(f)
        rcm = self._mols.weighted(r)
        W = self.get_atomic_virial().value_in_unit(unit.kilojoules_per_mole)
        return (W + float(np.sum(rcm*fcm)) - float(np.sum(r*f)))*unit.kilojoules_per_mole

    def import_configuration(self, state):
        self.setPeriodicBoxVectors(*state.getPeriodicBoxVectors())
        self.setPositions(state.getPositions(asNumpy=True))
        self.setVelocities(state.getVelocities(asNumpy=True))
        self._make_obsolete()

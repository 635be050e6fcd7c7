"""
Application layer: PDB + force-field XML loading and the ``Simulation`` convenience object.

The reference's tests build every input through ``simtk.openmm.app`` (PDBFile, ForceField,
Simulation; e.g. tests/test_respa_forces.py:15-23).  OpenMM's app layer is outside the hot
path but it defines the inputs, so the subset those tests use is restated here: fixed-column
PDB reading, residue templates matched by atom name, harmonic bond / angle / periodic torsion
/ nonbonded generators, rigid-water and H-bond constraints, 1-2/1-3 exclusions and scaled 1-4
exceptions (OpenMM semantics: SURVEY appendix A5-A7, D3).
"""

import math
import xml.etree.ElementTree as etree

import numpy as np

from . import mm
from . import unit

NoCutoff = mm.NonbondedForce.NoCutoff
CutoffNonPeriodic = mm.NonbondedForce.CutoffNonPeriodic
CutoffPeriodic = mm.NonbondedForce.CutoffPeriodic
Ewald = mm.NonbondedForce.Ewald
PME = mm.NonbondedForce.PME


class _Constraint(object):
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return self.name


HBonds = _Constraint('HBonds')
AllBonds = _Constraint('AllBonds')
HAngles = _Constraint('HAngles')


class Atom(object):
    def __init__(self, name, element, index, residue):
        self.name, self.element, self.index, self.residue = name, element, index, residue
        self.id = str(index + 1)


class Residue(object):
    def __init__(self, name, index, chain):
        self.name, self.index, self.chain = name, index, chain
        self._atoms = []

    def atoms(self):
        return iter(self._atoms)


class Topology(object):
    def __init__(self):
        self._atoms = []
        self._residues = []
        self._bonds = []
        self._box = None

    def addResidue(self, name, chain=None):
        residue = Residue(name, len(self._residues), chain)
        self._residues.append(residue)
        return residue

    def addAtom(self, name, element, residue):
        atom = Atom(name, element, len(self._atoms), residue)
        self._atoms.append(atom)
        residue._atoms.append(atom)
        return atom

    def addBond(self, atom1, atom2):
        self._bonds.append((atom1, atom2))

    def atoms(self):
        return iter(self._atoms)

    def residues(self):
        return iter(self._residues)

    def bonds(self):
        return iter(self._bonds)

    def getNumAtoms(self):
        return len(self._atoms)

    def getNumResidues(self):
        return len(self._residues)

    def getPeriodicBoxVectors(self):
        return self._box

    def setPeriodicBoxVectors(self, vectors):
        self._box = vectors

    def getUnitCellDimensions(self):
        if self._box is None:
            return None
        v = unit.md_value(self._box)
        return mm.Vec3(v[0][0], v[1][1], v[2][2])*unit.nanometer


class PDBFile(object):
    """Fixed-column PDB reader (both dialects in the reference's tests/data)."""

    def __init__(self, file):
        top = self.topology = Topology()
        coords = []
        residue = None
        key = None
        force_new = True
        opened = isinstance(file, str)
        handle = open(file) if opened else file
        try:
            for line in handle:
                record = line[:6].strip()
                if record == 'CRYST1':
                    a, b, c = (float(line[6:15]), float(line[15:24]), float(line[24:33]))
                    angles = [float(line[33:40]), float(line[40:47]), float(line[47:54])]
                    if any(abs(x - 90.0) > 1e-6 for x in angles):
                        raise ValueError('only orthorhombic boxes are supported')
                    top.setPeriodicBoxVectors([mm.Vec3(0.1*a, 0, 0), mm.Vec3(0, 0.1*b, 0),
                                               mm.Vec3(0, 0, 0.1*c)]*unit.nanometer)
                elif record in ('ATOM', 'HETATM'):
                    name = line[12:16].strip()
                    resname = line[17:21].strip()
                    newkey = (resname, line[21], line[22:27].strip())
                    if force_new or newkey != key:
                        residue = top.addResidue(resname, line[21])
                        key, force_new = newkey, False
                    element = line[76:78].strip() if len(line) > 76 else ''
                    top.addAtom(name, element or name[0], residue)
                    coords.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
                elif record == 'TER':
                    force_new = True
                elif record == 'ENDMDL':
                    break
        finally:
            if opened:
                handle.close()
        self._positions = 0.1*np.array(coords, dtype=np.float64)
        self.positions = unit.Quantity([mm.Vec3(*row) for row in self._positions], unit.nanometer)

    def getPositions(self, asNumpy=False):
        if asNumpy:
            return unit.Quantity(self._positions.copy(), unit.nanometer)
        return self.positions

    def getTopology(self):
        return self.topology


def _is_water(residue):
    atoms = residue._atoms
    if len(atoms) != 3:
        return False
    elements = sorted(a.element.upper()[:1] for a in atoms)
    return elements == ['H', 'H', 'O']


class ForceField(object):
    def __init__(self, *files):
        self._types = {}          # name -> (class, element, mass)
        self._templates = {}      # residue name -> dict(atoms=[(name,type,charge)], bonds=[(n1,n2)])
        self._bonds = []          # (t1, t2, length, k)   t* = set of matching type names
        self._angles = []
        self._propers = []        # (t1..t4, [(n, phase, k)])
        self._impropers = []
        self._nb_params = {}      # type -> (charge or None, sigma, eps)
        self._c14 = 1.0
        self._lj14 = 1.0
        self._has = set()
        for f in files:
            self._load(f)

    def _match_set(self, attrib, index):
        if 'type%d' % index in attrib:
            t = attrib['type%d' % index]
            return None if t == '' else {t}
        c = attrib.get('class%d' % index, '')
        if c == '':
            return None
        return {name for name, (cls, _, _) in self._types.items() if cls == c}

    def _load(self, filename):
        root = etree.parse(filename).getroot()
        for node in root.findall('AtomTypes/Type'):
            a = node.attrib
            self._types[a['name']] = (a.get('class', a['name']), a.get('element', ''), float(a['mass']))
        for res in root.findall('Residues/Residue'):
            atoms = [(n.attrib['name'], n.attrib['type'], float(n.attrib.get('charge', 'nan')))
                     for n in res.findall('Atom')]
            names = [a[0] for a in atoms]
            bonds = []
            for b in res.findall('Bond'):
                if 'atomName1' in b.attrib:
                    bonds.append((b.attrib['atomName1'], b.attrib['atomName2']))
                else:
                    bonds.append((names[int(b.attrib['from'])], names[int(b.attrib['to'])]))
            self._templates[res.attrib['name']] = dict(atoms=atoms, bonds=bonds)
        node = root.find('HarmonicBondForce')
        if node is not None:
            self._has.add('bond')
            for b in node.findall('Bond'):
                self._bonds.append((self._match_set(b.attrib, 1), self._match_set(b.attrib, 2),
                                    float(b.attrib['length']), float(b.attrib['k'])))
        node = root.find('HarmonicAngleForce')
        if node is not None:
            self._has.add('angle')
            for a in node.findall('Angle'):
                self._angles.append((self._match_set(a.attrib, 1), self._match_set(a.attrib, 2),
                                     self._match_set(a.attrib, 3), float(a.attrib['angle']), float(a.attrib['k'])))
        node = root.find('PeriodicTorsionForce')
        if node is not None:
            self._has.add('torsion')
            for kind, store in (('Proper', self._propers), ('Improper', self._impropers)):
                for t in node.findall(kind):
                    terms = []
                    n = 1
                    while 'periodicity%d' % n in t.attrib:
                        terms.append((int(t.attrib['periodicity%d' % n]), float(t.attrib['phase%d' % n]),
                                      float(t.attrib['k%d' % n])))
                        n += 1
                    store.append(tuple(self._match_set(t.attrib, i) for i in (1, 2, 3, 4)) + (terms,))
        node = root.find('NonbondedForce')
        if node is not None:
            self._has.add('nonbonded')
            self._c14 = float(node.attrib.get('coulomb14scale', 1.0))
            self._lj14 = float(node.attrib.get('lj14scale', 1.0))
            from_residue = {n.attrib['name'] for n in node.findall('UseAttributeFromResidue')}
            for a in node.findall('Atom'):
                types = {a.attrib['type']} if 'type' in a.attrib else \
                    {name for name, (cls, _, _) in self._types.items() if cls == a.attrib['class']}
                charge = None if 'charge' in from_residue else float(a.attrib['charge'])
                for t in types:
                    self._nb_params[t] = (charge, float(a.attrib['sigma']), float(a.attrib['epsilon']))

    # -- plain-data round trip (used for test fixtures that must travel without the XML files) -----
    def to_dict(self):
        def sets(t):
            return [None if x is None else sorted(x) for x in t]
        return dict(types={k: list(v) for k, v in self._types.items()},
                    templates=self._templates,
                    bonds=[sets(b[:2]) + list(b[2:]) for b in self._bonds],
                    angles=[sets(a[:3]) + list(a[3:]) for a in self._angles],
                    propers=[sets(p[:4]) + [p[4]] for p in self._propers],
                    nb_params={k: list(v) for k, v in self._nb_params.items()},
                    c14=self._c14, lj14=self._lj14, has=sorted(self._has))

    @classmethod
    def from_dict(cls, data):
        def sets(t):
            return tuple(None if x is None else set(x) for x in t)
        self = cls()
        self._types = {k: tuple(v) for k, v in data['types'].items()}
        self._templates = {k: dict(atoms=[tuple(a) for a in v['atoms']], bonds=[tuple(b) for b in v['bonds']])
                           for k, v in data['templates'].items()}
        self._bonds = [sets(b[:2]) + tuple(b[2:]) for b in data['bonds']]
        self._angles = [sets(a[:3]) + tuple(a[3:]) for a in data['angles']]
        self._propers = [sets(p[:4]) + ([tuple(t) for t in p[4]],) for p in data['propers']]
        self._nb_params = {k: tuple(v) for k, v in data['nb_params'].items()}
        self._c14, self._lj14, self._has = data['c14'], data['lj14'], set(data['has'])
        return self

    @staticmethod
    def _fits(pattern, value):
        return pattern is None or value in pattern

    def createSystem(self, topology, nonbondedMethod=NoCutoff, nonbondedCutoff=1.0*unit.nanometer,
                     constraints=None, rigidWater=True, removeCMMotion=True, hydrogenMass=None,
                     ewaldErrorTolerance=0.0005, useDispersionCorrection=True, **kwargs):
        system = mm.System()
        atoms = list(topology.atoms())
        types = [None]*len(atoms)
        charges = [0.0]*len(atoms)
        bonds = []
        for residue in topology.residues():
            template = self._templates.get(residue.name)
            res_atoms = residue._atoms
            if template is None or len(template['atoms']) != len(res_atoms):
                names = sorted(a.name for a in res_atoms)
                candidates = [t for t in self._templates.values()
                              if sorted(x[0] for x in t['atoms']) == names]
                if not candidates:
                    raise ValueError('No template found for residue %d (%s)' % (residue.index + 1, residue.name))
                template = candidates[0]
            by_name = {a.name: a for a in res_atoms}
            for name, type_name, charge in template['atoms']:
                if name not in by_name:
                    raise ValueError('atom %s missing in residue %s' % (name, residue.name))
                types[by_name[name].index] = type_name
                charges[by_name[name].index] = charge
            for n1, n2 in template['bonds']:
                bonds.append((by_name[n1].index, by_name[n2].index))
        existing = set((a.index, b.index) for a, b in topology.bonds())
        for a, b in bonds:
            if (a, b) not in existing and (b, a) not in existing:
                topology.addBond(atoms[a], atoms[b])
        for i, atom in enumerate(atoms):
            system.addParticle(self._types[types[i]][2])
        box = topology.getPeriodicBoxVectors()
        if box is not None:
            system.setDefaultPeriodicBoxVectors(*unit.md_value(box))

        is_h = [self._types[t][1].upper() == 'H' or atoms[i].element.upper() == 'H' for i, t in enumerate(types)]
        water = [_is_water(a.residue) for a in atoms]
        neighbours = [[] for _ in atoms]
        for a, b in bonds:
            neighbours[a].append(b)
            neighbours[b].append(a)

        # --- bonds ---------------------------------------------------------------------
        constrained_length = {}
        bond_force = mm.HarmonicBondForce()
        for a, b in bonds:
            params = None
            for t1, t2, length, k in self._bonds:
                if (self._fits(t1, types[a]) and self._fits(t2, types[b])) or \
                   (self._fits(t1, types[b]) and self._fits(t2, types[a])):
                    params = (length, k)
                    break
            if params is None:
                continue
            constrain = (rigidWater and water[a]) or constraints in (AllBonds, HAngles) or \
                (constraints is HBonds and (is_h[a] or is_h[b]))
            if constrain:
                system.addConstraint(a, b, params[0])
                constrained_length[(min(a, b), max(a, b))] = params[0]
            else:
                bond_force.addBond(a, b, params[0], params[1])
        if 'bond' in self._has:
            system.addForce(bond_force)

        # --- angles --------------------------------------------------------------------
        angle_force = mm.HarmonicAngleForce()
        angle_list = []
        for j in range(len(atoms)):
            nb = neighbours[j]
            for x in range(len(nb)):
                for y in range(x + 1, len(nb)):
                    angle_list.append((nb[x], j, nb[y]))
        for i, j, k in angle_list:
            params = None
            for t1, t2, t3, theta, K in self._angles:
                if self._fits(t2, types[j]) and (
                        (self._fits(t1, types[i]) and self._fits(t3, types[k])) or
                        (self._fits(t1, types[k]) and self._fits(t3, types[i]))):
                    params = (theta, K)
                    break
            if params is None:
                continue
            constrain = (rigidWater and water[j]) or \
                (constraints is HAngles and ((is_h[i] and is_h[k]) or
                                             ((is_h[i] or is_h[k]) and atoms[j].element.upper() == 'O')))
            if constrain:
                l1 = constrained_length.get((min(i, j), max(i, j)))
                l2 = constrained_length.get((min(k, j), max(k, j)))
                if l1 is not None and l2 is not None:
                    length = math.sqrt(l1*l1 + l2*l2 - 2*l1*l2*math.cos(params[0]))
                    system.addConstraint(i, k, length)
                    continue
            angle_force.addAngle(i, j, k, params[0], params[1])
        if 'angle' in self._has:
            system.addForce(angle_force)

        # --- proper torsions -----------------------------------------------------------
        if 'torsion' in self._has:
            torsion_force = mm.PeriodicTorsionForce()
            seen = set()
            for j, k in bonds:
                for i in neighbours[j]:
                    if i == k:
                        continue
                    for l in neighbours[k]:
                        if l == j or l == i:
                            continue
                        key = (i, j, k, l) if i < l else (l, k, j, i)
                        if key in seen:
                            continue
                        seen.add(key)
                        tt = [types[x] for x in (i, j, k, l)]
                        match = None
                        for t1, t2, t3, t4, terms in self._propers:
                            pats = (t1, t2, t3, t4)
                            fwd = all(self._fits(p, t) for p, t in zip(pats, tt))
                            rev = all(self._fits(p, t) for p, t in zip(pats, tt[::-1]))
                            if fwd or rev:
                                wild = sum(p is None for p in pats)
                                if match is None or wild < match[0]:
                                    match = (wild, terms)
                                    if wild == 0:
                                        break
                        if match is not None:
                            for n, phase, kk in match[1]:
                                if kk != 0:
                                    torsion_force.addTorsion(i, j, k, l, n, phase, kk)
            system.addForce(torsion_force)

        # --- nonbonded -----------------------------------------------------------------
        if 'nonbonded' in self._has:
            nb = mm.NonbondedForce()
            for i, t in enumerate(types):
                q, sigma, eps = self._nb_params[t]
                nb.addParticle(charges[i] if q is None else q, sigma, eps)
            nb.createExceptionsFromBonds(bonds, self._c14, self._lj14)
            nb.setNonbondedMethod(nonbondedMethod)
            nb.setCutoffDistance(nonbondedCutoff)
            nb.setEwaldErrorTolerance(ewaldErrorTolerance)
            nb.setUseDispersionCorrection(useDispersionCorrection)
            system.addForce(nb)
        if removeCMMotion:
            system.addForce(mm.CMMotionRemover())
        return system


class Simulation(object):
    def __init__(self, topology, system, integrator, platform=None, platformProperties=None):
        from . import engine
        self.topology = topology
        self.system = system
        self.integrator = integrator
        self.currentStep = 0
        self.reporters = []
        if platform is None:
            platform = mm.Platform.getPlatformByName('B200')
        self.context = engine.Context(system, integrator, platform, platformProperties or {})

    def step(self, steps):
        end = self.currentStep + steps
        while self.currentStep < end:
            chunk = end - self.currentStep
            for reporter in self.reporters:
                due = reporter.describeNextReport(self)[0]
                chunk = min(chunk, due)
            self.integrator.step(chunk)
            self.currentStep += chunk
            for reporter in self.reporters:
                info = reporter.describeNextReport(self)
                if info[0] == 0 or (self.currentStep % getattr(reporter, '_reportInterval', 1)) == 0:
                    state = self.context.getState(getPositions=info[1], getVelocities=info[2],
                                                  getForces=info[3], getEnergy=info[4])
                    reporter.report(self, state)

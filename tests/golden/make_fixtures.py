"""
Convert the reference's test inputs (tests/data/*.pdb + *.xml under /root/reference) into
self-contained fixtures tests/golden/systems/<case>.json.gz: box, residues, atom names, elements,
positions and the parsed force field.  Runs only in the build container.

    python tests/golden/make_fixtures.py
"""

import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
DATA = '/root/reference/tests/data'
CASES = ['q-SPC-FW', 'emim_BCN4_Jiung2014', 'methane-in-water', 'phenol-in-water',
         'hydroxyethylaminoanthraquinone-in-water']


def main():
    from atomsmm_b200 import app, unit
    out_dir = os.path.join(ROOT, 'tests', 'golden', 'systems')
    os.makedirs(out_dir, exist_ok=True)
    for case in CASES:
        pdb = app.PDBFile(os.path.join(DATA, case + '.pdb'))
        ff = app.ForceField(os.path.join(DATA, case + '.xml'))
        box = pdb.topology.getUnitCellDimensions().value_in_unit(unit.nanometer)
        residues = [[r.name, r.chain, [a.name for a in r.atoms()], [a.element for a in r.atoms()]]
                    for r in pdb.topology.residues()]
        positions = pdb.getPositions(asNumpy=True).value_in_unit(unit.nanometer)
        record = dict(case=case, box=list(box), residues=residues,
                      positions=[[repr(float(v)) for v in row] for row in positions], forcefield=ff.to_dict())
        path = os.path.join(out_dir, case + '.json.gz')
        with gzip.GzipFile(path, 'wb', mtime=0) as handle:
            handle.write(json.dumps(record, sort_keys=True).encode())
        print(path, os.path.getsize(path))


if __name__ == '__main__':
    main()

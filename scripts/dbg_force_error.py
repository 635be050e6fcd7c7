import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import atomsmm_b200 as atomsmm
from atomsmm_b200 import app, mm, unit
from oracle import refmath, cport
import systems
system, pdb = systems.flexible('q-SPC-FW', app.CutoffPeriodic)
respa = atomsmm.RESPASystem(system, 7*systems.A, 5*systems.A)
big, pos = systems.replicate(respa, systems.positions_of(pdb), refmath.system_box(system), 2)
context = mm.Context(big, mm.VerletIntegrator(0.0), mm.Platform.getPlatformByName('B200'))
context.setPositions(pos)
port = cport.CPort(big)
for groups in ({1}, {2}, {0}):
    state = context.getState(getEnergy=True, getForces=True, groups=groups)
    f = state.getForces(asNumpy=True).value_in_unit(unit.kilojoules_per_mole/unit.nanometer)
    fr, e, w = port.evaluate(pos, groups)
    err = np.sqrt(((f-fr)**2).sum(1)); mag=np.sqrt((fr**2).sum(1))
    print(groups, 'rms', np.sqrt(((f-fr)**2).sum()/(fr**2).sum()), 'max abs err', err.max(), 'at', err.argmax(), 'mag', mag[err.argmax()], 'n>0.05', (err>0.05).sum(), 'median err', np.median(err), 'rms force', np.sqrt((fr**2).sum(1).mean()))
    idx=np.argsort(err)[-5:]
    print('  worst', idx, err[idx], f[idx[-1]], fr[idx[-1]])

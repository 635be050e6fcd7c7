"""NVE drift probe: config 1 (1 536-atom RESPASystem water), RespaPropagator(loops) at dt, GPU engine beside the
float64 C oracle; prints total energies and the fitted slopes.  Usage: drift_probe.py [steps] [dt_fs] [n0 n1]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402
import atomsmm_b200 as atomsmm  # noqa: E402
from atomsmm_b200 import mm, unit  # noqa: E402
import systems  # noqa: E402
from test_gpu_integrators import thermal_velocities  # noqa: E402
from oracle import cport  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dt_fs = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
n0, n1 = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (4, 2)
respa, pdb = systems.respa_water()
pos = systems.positions_of(pdb)
vel = thermal_velocities(respa, 300.0, 1234)
dof = atomsmm.countDegreesOfFreedom(respa)
mass = np.array(respa._masses)
integrator = atomsmm.RespaPropagator([n0, n1, 1]).integrator(dt_fs*unit.femtoseconds)
props = {}
if os.environ.get('B2_SKIN'):
    props['Skin'] = float(os.environ['B2_SKIN'])
context = mm.Context(respa, integrator, mm.Platform.getPlatformByName('B200'), props)
context.setPositions(pos)
context.setVelocities(vel)
port = cport.CPort(respa)
x, v = pos.copy(), vel.copy()
gpu, cpu, cross = [], [], []
block = 20
for k in range(steps//block):
    integrator.step(block)
    s = context.getState(getEnergy=True, getPositions=True, getVelocities=True, groups={0, 2})
    gpu.append(s._potential + s._kinetic)
    # the GPU configuration evaluated by the oracle: separates "energy evaluation" from "trajectory" effects
    cross.append(port.evaluate(s._positions, {0, 2})[1] + 0.5*float(np.sum(mass[:, None]*s._velocities**2)))
    if os.environ.get('B2_NO_ORACLE_TRAJ') is None:
        x, v, _ = port.respa(x, v, block, dt_fs*1e-3, n0, n1)
        cpu.append(port.evaluate(x, {0, 2})[1] + 0.5*float(np.sum(mass[:, None]*v*v)))
t = np.arange(1, len(gpu) + 1)*block*dt_fs*1e-3
print('dof', dof, 'kT', 2.494)
print('gpu  slope %.5f kJ/mol/ps/dof  (E0 %.3f, E_end %.3f, std of residual %.3f)' % (
    np.polyfit(t, gpu, 1)[0]/dof, gpu[0], gpu[-1], np.std(np.array(gpu) - np.polyval(np.polyfit(t, gpu, 1), t))))
print('gpu* slope %.5f (GPU trajectory, oracle energies; max |E_gpu - E_oracle| %.4f)' % (
    np.polyfit(t, cross, 1)[0]/dof, np.max(np.abs(np.array(gpu) - np.array(cross)))))
if cpu:
    print('cpu  slope %.5f kJ/mol/ps/dof  (E0 %.3f, E_end %.3f, std of residual %.3f)' % (
        np.polyfit(t, cpu, 1)[0]/dof, cpu[0], cpu[-1], np.std(np.array(cpu) - np.polyval(np.polyfit(t, cpu, 1), t))))
print('gpu', ' '.join('%.2f' % e for e in gpu[::5]))
if cpu:
    print('cpu', ' '.join('%.2f' % e for e in cpu[::5]))
print('rebuilds', context.list_stats())

"""Where does an end-to-end bench step (host buffers in, host buffers out) spend its host time?"""
import cProfile
import os
import pstats
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from atomsmm_b200 import mm, unit  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
system, pos, vel = bench.build_workload(reps)
integrator, dof = bench.make_integrator(system)
context = mm.Context(system, integrator, mm.Platform.getPlatformByName('B200'))
context.setPositions(pos)
context.setVelocities(vel)
integrator.step(1300 if reps <= 4 else 100)
host_x = torch.from_numpy(pos.copy()).pin_memory()
host_v = torch.from_numpy(vel.copy()).pin_memory()


def e2e_step(md=100):
    context.setPositions(host_x)
    context.setVelocities(host_v)
    integrator.step(md)
    s = context.getState(getPositions=True, getVelocities=True, getEnergy=True)
    host_x.copy_(torch.from_numpy(s._positions))
    host_v.copy_(torch.from_numpy(s._velocities))
    return s._potential + s._kinetic


s = context.getState(getPositions=True, getVelocities=True)
host_x.copy_(torch.from_numpy(s._positions))
host_v.copy_(torch.from_numpy(s._velocities))
e2e_step()
for name, fn in (('setPositions', lambda: context.setPositions(host_x)), ('setVelocities', lambda: context.setVelocities(host_v)),
                 ('step100', lambda: (integrator.step(100), context.synchronize())),
                 ('getState', lambda: context.getState(getPositions=True, getVelocities=True, getEnergy=True))):
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    print('%-14s %8.2f ms' % (name, (time.perf_counter() - t0)/3*1e3))
for k in range(6):
    t0 = time.perf_counter(); context.setPositions(host_x); context.synchronize()
    t1 = time.perf_counter(); context.setVelocities(host_v)
    t2 = time.perf_counter(); integrator.step(100); context.synchronize()
    t3 = time.perf_counter(); s = context.getState(getPositions=True, getVelocities=True, getEnergy=True)
    t4 = time.perf_counter(); host_x.copy_(torch.from_numpy(s._positions)); host_v.copy_(torch.from_numpy(s._velocities))
    t5 = time.perf_counter()
    print('e2e %d: setPos %.1f setVel %.1f step %.1f getState %.1f copy %.1f ms  %s' % (
        k, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3, context.list_stats()))
prof = cProfile.Profile()
prof.enable()
for _ in range(3):
    e2e_step()
prof.disable()
pstats.Stats(prof).sort_stats('cumulative').print_stats(18)

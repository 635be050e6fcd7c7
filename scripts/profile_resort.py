"""Cost of the re-ordering path of setPositions (triggered by a rigid 0.5 nm shift of the box contents)."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import numpy as np  # noqa: E402
from atomsmm_b200 import mm  # noqa: E402

system, pos, vel = bench.build_workload(4)
integrator, dof = bench.make_integrator(system)
context = mm.Context(system, integrator, mm.Platform.getPlatformByName('B200'))
context.setPositions(pos)
context.setVelocities(vel)
integrator.step(100)
context.synchronize()
for k in range(4):
    shifted = pos + 0.5*(k + 1)
    t0 = time.perf_counter()
    context.setPositions(shifted)
    context.synchronize()
    t1 = time.perf_counter()
    integrator.step(1)
    context.synchronize()
    t2 = time.perf_counter()
    integrator.step(1)
    context.synchronize()
    t3 = time.perf_counter()
    integrator.step(100)
    context.synchronize()
    t4 = time.perf_counter()
    print('resort setPositions %.1f ms, first step %.1f ms, second step %.1f ms, 100 steps %.1f ms' % (
        (t1 - t0)*1e3, (t2 - t1)*1e3, (t3 - t2)*1e3, (t4 - t3)*1e3))

"""Small driver for ncu: build the C2 workload, run a few MD steps."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from atomsmm_b200 import mm  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
system, pos, vel = bench.build_workload(reps)
integrator, dof = bench.make_integrator(system)
context = mm.Context(system, integrator, mm.Platform.getPlatformByName('B200'))
context.setPositions(pos)
context.setVelocities(vel)
integrator.step(steps)
context.synchronize()
print('done', context.counters())

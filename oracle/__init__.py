"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

A float64 CPU restatement of the arithmetic the reference (atoms-ufrj/atomsmm) delegates to the
OpenMM *Reference* platform for the hot path: pair potentials given as algebraic strings,
bonded terms, PME, the virial-as-energy system of PressureComputer and the CustomIntegrator step
programs.  OpenMM itself is a third-party dependency that is neither vendored under
/root/reference nor pinned (ci/environment.yml:7 ``- openmm``; era 7.2-7.5 by the ``simtk``
namespace, forces.py:18), so its published algorithm is restated here and anchored on the
reference's own known-answer tests (tests/test_respa_forces.py, test_DampedSmoothedForce.py,
test_ExceptionNonbondedForce.py, test_computers.py, test_systems.py:131-152).

Parity status: single-point energies / virials / pressures are PINNED by those goldens
(tests/test_oracle_goldens.py).  Integrator trajectories are *parity unpinned*: the reference's
only trajectory goldens (tests/test_propagators.py) depend on OpenMM's SFMT random stream and
SHAKE/CCMA and cannot be reproduced without OpenMM; the step-program interpreter is validated
structurally (programs captured from the reference's Python layer) and by invariants.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (atomsmm_b200/) never does.
"""

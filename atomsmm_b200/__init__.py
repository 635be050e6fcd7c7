"""
atomsmm_b200 -- a B200-native engine behind the atomsmm API for its one hot path: near/far-split
nonbonded pair forces, their virial, and RESPA / thermostat integrator updates.

The public names mirror ``atomsmm.__init__`` (reference: src/atomsmm/__init__.py:3-43) for the
hot-path subset; ``mm`` / ``app`` / ``unit`` stand in for the
``simtk.openmm`` / ``simtk.openmm.app`` / ``simtk.unit`` modules the reference's users import.
"""

__version__ = '0.1.0'

from . import unit  # noqa: F401
from . import mm  # noqa: F401
from . import app  # noqa: F401
from . import forces  # noqa: F401
from . import propagators  # noqa: F401
from . import integrators  # noqa: F401
from . import systems  # noqa: F401
from . import utils  # noqa: F401
from .forces import DampedSmoothedForce  # noqa: F401
from .forces import FarNonbondedForce  # noqa: F401
from .forces import NearExceptionForce  # noqa: F401
from .forces import NearNonbondedForce  # noqa: F401
from .forces import NonbondedExceptionsForce  # noqa: F401
from .forces import SoftcoreForce  # noqa: F401
from .forces import SoftcoreLennardJonesForce  # noqa: F401
from .integrators import AdiabaticDynamicsIntegrator  # noqa: F401
from .integrators import ExtendedSystemVariable  # noqa: F401
from .integrators import GlobalThermostatIntegrator  # noqa: F401
from .integrators import Langevin_R_Integrator  # noqa: F401
from .integrators import MultipleTimeScaleIntegrator  # noqa: F401
from .integrators import NHL_R_Integrator  # noqa: F401
from .propagators import ChainedPropagator  # noqa: F401
from .propagators import GenericBoostPropagator  # noqa: F401
from .propagators import NoseHooverChainPropagator  # noqa: F401
from .propagators import NoseHooverLangevinPropagator  # noqa: F401
from .propagators import NoseHooverPropagator  # noqa: F401
from .propagators import OrnsteinUhlenbeckPropagator  # noqa: F401
from .propagators import RespaPropagator  # noqa: F401
from .propagators import SplitPropagator  # noqa: F401
from .propagators import SuzukiYoshidaPropagator  # noqa: F401
from .propagators import TranslationPropagator  # noqa: F401
from .propagators import TrotterSuzukiPropagator  # noqa: F401
from .propagators import VelocityBoostPropagator  # noqa: F401
from .propagators import VelocityRescalingPropagator  # noqa: F401
from .propagators import VelocityVerletPropagator  # noqa: F401
from .systems import AlchemicalSystem  # noqa: F401
from .systems import ComputingSystem  # noqa: F401
from .systems import RESPASystem  # noqa: F401
from .utils import InputError  # noqa: F401
from .utils import countDegreesOfFreedom  # noqa: F401
from .utils import evaluateForce  # noqa: F401
from .utils import findNonbondedForce  # noqa: F401
from .utils import hijackForce  # noqa: F401
from .utils import splitPotentialEnergy  # noqa: F401


def __getattr__(name):
    if name == 'PressureComputer':
        from .computers import PressureComputer
        return PressureComputer
    if name == 'computers':
        import importlib
        return importlib.import_module('.computers', __name__)
    raise AttributeError(name)

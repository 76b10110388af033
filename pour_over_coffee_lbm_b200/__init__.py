"""pour_over_coffee_lbm_b200 -- B200-native (sm_100a) D3Q19 lattice-Boltzmann hot path of
latteine1217/pour-over-coffee-lbm, behind the reference's solver API.

Host = Python (ctypes over a C-ABI shared library, torch tensors for device memory);
compute = hand-written CUDA in csrc/.  No Triton, no multi-backend dispatch, no CPU fallback.
"""
from . import _lib
from .config import LBMConfig
from .errors import (BackendError, BackendInitializationError, ComputeExecutionError,
                     MemoryAllocationError, PerformanceDegradationError, PlatformDetectionError)

__all__ = ["_lib", "LBMConfig", "BackendError", "BackendInitializationError", "ComputeExecutionError",
           "MemoryAllocationError", "PerformanceDegradationError", "PlatformDetectionError"]


def __getattr__(name):
    # heavy modules (torch) are imported lazily so that `import pour_over_coffee_lbm_b200` stays cheap
    if name in ("D3Q19Engine", "ParticleState", "particles_couple"):
        from . import engine
        return getattr(engine, name)
    if name in ("LBMSolver", "UnifiedLBMSolver", "create_unified_solver"):
        from . import solver
        return getattr(solver, name)
    if name in ("B200Backend", "ComputeBackend"):
        from . import backend
        return getattr(backend, name)
    raise AttributeError(name)

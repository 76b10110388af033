#!/usr/bin/env python
"""Records the reference's API surface for the hot path (run in the authoring container): the members of
LBMSolverProtocol, the public methods of the classes the facade mirrors, the ComputeBackend interface and the error
hierarchy -- read by IMPORTING the reference's modules (under tests/golden/taichi_shim).  Written to
tests/golden/reference_api_surface.json; tests/test_host_logic.py checks the facade classes against it."""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_reference_goldens import load_reference, quiet  # noqa: E402


def public_methods(cls):
    return sorted(n for n, v in vars(cls).items() if not n.startswith("_") and (inspect.isfunction(v) or callable(v)))


if __name__ == "__main__":
    load_reference(16)
    with quiet():
        from src.core.lbm_protocol import LBMSolverProtocol
        from src.core.legacy.lbm_solver import LBMSolver
        from src.core.backends import compute_backends as CB
        from src.physics.filter_paper import FilterPaperSystem
        from src.physics.pressure_gradient_drive import PressureGradientDrive
        from src.physics.coffee_particles import CoffeeParticleSystem
        from src.physics.boundary_conditions import BoundaryConditionManager
        from src.physics.les_turbulence import LESTurbulenceModel
        from src.core.multiphase_3d import MultiphaseFlow3D
        from src.physics.precise_pouring import PrecisePouringSystem
        s = LBMSolver()
    proto = sorted(getattr(LBMSolverProtocol, "__protocol_attrs__", set()))
    out = {
        "LBMSolverProtocol": proto,
        "LBMSolver.methods": public_methods(LBMSolver),
        "LBMSolver.instance_fields": sorted(n for n, v in vars(s).items() if not n.startswith("_") and type(v).__name__ in ("Field", "VectorField")),
        "ComputeBackend.abstract": sorted(getattr(CB.ComputeBackend, "__abstractmethods__", [])),
        "ComputeBackend.methods": public_methods(CB.ComputeBackend),
        "errors": sorted(n for n, v in vars(CB).items() if inspect.isclass(v) and issubclass(v, Exception)),
        "FilterPaperSystem.methods": public_methods(FilterPaperSystem),
        "PressureGradientDrive.methods": public_methods(PressureGradientDrive),
        "CoffeeParticleSystem.methods": public_methods(CoffeeParticleSystem),
        "BoundaryConditionManager.methods": public_methods(BoundaryConditionManager),
        "LESTurbulenceModel.methods": public_methods(LESTurbulenceModel),
        "MultiphaseFlow3D.methods": public_methods(MultiphaseFlow3D),
        "PrecisePouringSystem.methods": public_methods(PrecisePouringSystem),
    }
    with open(os.path.join(HERE, "reference_api_surface.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, len(v))

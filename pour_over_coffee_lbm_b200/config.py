"""Constants of the hot path, mirroring the reference's `config` package.

Same names and values as config/core.py:23-93 and config/physics.py:20-190 of the reference
(evaluated once at import there; here `LBMConfig` evaluates the same formulas for any grid so
the 256^3 / 512^3 / 1024^3 BASELINE configs can be expressed).  Only what the D3Q19 step, the
V60 geometry and the particle coupling read is kept.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

Q_3D = 19
# config/core.py:36-47
CX_3D = np.array([0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0], dtype=np.int32)
CY_3D = np.array([0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1], dtype=np.int32)
CZ_3D = np.array([0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1], dtype=np.int32)
WEIGHTS_3D = np.array([1.0 / 3.0] + [1.0 / 18.0] * 6 + [1.0 / 36.0] * 12, dtype=np.float32)
OPPOSITE_3D = np.array([0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15], dtype=np.int32)


@dataclass
class LBMConfig:
    NX: int = 224
    NY: int = 224
    NZ: int = 224
    DX: float = 1.0
    DT: float = 1.0
    CS2: float = 1.0 / 3.0
    INV_CS2: float = 3.0
    TAU_FLUID: float = 0.53
    TAU_AIR: float = 0.8
    RHO_0: float = 1.0
    SCALE_VELOCITY: float = 0.01
    TIME_SCALE_OPTIMIZATION_FACTOR: float = 1.2
    PHYSICAL_DOMAIN_SIZE: float = 0.14
    SMAGORINSKY_CONSTANT: float = 0.17       # config/core.py:90 (unused by the reference's LES)
    LES_CS: float = 0.18                     # les_turbulence.py:95 (the value actually used)
    ENABLE_LES: bool = True
    LES_REYNOLDS_THRESHOLD: float = 500.0
    GRAVITY_PHYS: float = 9.81
    GRAVITY_STRENGTH_FACTOR: float = 0.5
    WATER_DENSITY_90C: float = 965.3
    WATER_VISCOSITY_90C: float = 3.15e-7
    COFFEE_BEAN_DENSITY: float = 1200.0
    CUP_HEIGHT: float = 0.085
    TOP_RADIUS: float = 0.058
    BOTTOM_RADIUS: float = 0.010
    PARTICLE_DIAMETER_MM: float = 0.65
    COFFEE_PARTICLE_RADIUS: float = 3.25e-4
    U_CHAR: float = 0.02
    PAPER_THICKNESS: float = 0.0001          # filter_paper.py:106
    PAPER_POROSITY: float = 0.85             # filter_paper.py:107
    AIR_DENSITY_20C: float = 1.204           # config/physics.py:50
    RHO_WATER: float = 1.0                   # config/physics.py:114
    WEBER_NUMBER: float = 1.0                # config/physics.py:180
    POUR_RATE_ML_S: float = 4.0              # config/physics.py:81
    POUR_HEIGHT_CM: float = 12.5             # config/physics.py:87
    INLET_DIAMETER_M: float = 0.005          # config/physics.py:88
    NOZZLE_DIAMETER_M: float = 0.005         # config/physics.py:89
    GRAVITY_CORRECTION: float = 0.05         # config/physics.py:92
    SCALE_LENGTH: float = field(default=0.0)
    SCALE_TIME: float = field(default=0.0)
    GRAVITY_LU: float = field(default=-1.0)

    def __post_init__(self):
        if self.SCALE_LENGTH == 0.0:
            self.SCALE_LENGTH = self.PHYSICAL_DOMAIN_SIZE / self.NZ
        if self.SCALE_TIME == 0.0:
            self.SCALE_TIME = (self.SCALE_LENGTH / self.SCALE_VELOCITY) * self.TIME_SCALE_OPTIMIZATION_FACTOR
        if self.GRAVITY_LU < 0.0:
            self.GRAVITY_LU = self.GRAVITY_PHYS * (self.SCALE_TIME ** 2) / self.SCALE_LENGTH * self.GRAVITY_STRENGTH_FACTOR

    TAU_WATER = property(lambda self: self.TAU_FLUID)            # config/__init__.py:285
    RE_CHAR = property(lambda self: self.U_CHAR * self.CUP_HEIGHT / self.WATER_VISCOSITY_90C)
    use_les = property(lambda self: self.ENABLE_LES and self.RE_CHAR > self.LES_REYNOLDS_THRESHOLD)

    # multiphase / pouring constants (config/physics.py:95-101, 115, 181, 254-273; config/core.py:84)
    RHO_AIR = property(lambda self: self.AIR_DENSITY_20C / self.WATER_DENSITY_90C)
    SURFACE_TENSION_LU = property(lambda self: (self.RHO_WATER * (self.U_CHAR * self.SCALE_TIME / self.SCALE_LENGTH) ** 2 * self.SCALE_LENGTH)
                                  / self.WEBER_NUMBER)
    GRID_SIZE_CM = property(lambda self: self.SCALE_LENGTH * 100)
    INLET_AREA = property(lambda self: math.pi * (min(self.INLET_DIAMETER_M, self.NOZZLE_DIAMETER_M) / 2.0) ** 2)

    @property
    def INLET_VELOCITY(self) -> float:
        base = (self.POUR_RATE_ML_S * 1e-6) / (math.pi * (self.INLET_DIAMETER_M / 2.0) ** 2)
        return min(0.05, base * (1.0 + self.GRAVITY_CORRECTION) * self.SCALE_TIME / self.SCALE_LENGTH)

    # --- f32 constants exactly as the reference's kernels see them --------------------------
    def v60_geometry_constants(self):
        """Python-scope f64 folds of filter_paper.py:239-249, 318-333, rounded to f32."""
        f = np.float32
        thick = np.maximum(f(1.0), f(self.PAPER_THICKNESS / self.SCALE_LENGTH))
        return [float(f(self.TOP_RADIUS / self.SCALE_LENGTH)), float(f(self.BOTTOM_RADIUS / self.SCALE_LENGTH)),
                float(f(self.CUP_HEIGHT / self.SCALE_LENGTH)), float(f(0.002 / self.SCALE_LENGTH)), float(thick)]

    def forchheimer_parameters(self):
        """filter_paper.py:423-469: kernel-scope f32 arithmetic (integer powers are multiplies)."""
        f = np.float32
        dp = f(self.PARTICLE_DIAMETER_MM * 1e-3)
        p = f(self.PAPER_POROSITY)
        one_m = f(1.0) - p
        k_phys = ((dp * dp) * ((p * p) * p)) / (f(180.0) * (one_m * one_m))
        beta = (f(1.75) * one_m) / ((p * p) * p)
        k_lu = k_phys / f(self.SCALE_LENGTH ** 2)
        return float(k_lu), float(beta)

    def filter_constants(self):
        """filter_paper.py:514-520, 578-586: constant folds feeding the drag coefficient."""
        f = np.float32
        return (float(f(self.WATER_VISCOSITY_90C * self.SCALE_TIME / (self.SCALE_LENGTH ** 2))),
                float(f(self.WATER_DENSITY_90C * self.SCALE_TIME ** 2 / (self.SCALE_LENGTH ** 3))))


DEFAULT = LBMConfig()

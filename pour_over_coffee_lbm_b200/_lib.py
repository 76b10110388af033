"""ctypes binding of liblbm_b200.so (the C ABI declared in include/lbm_b200.h).

This is the stub a maintainer of the reference would add: the reference is
Python, so the FFI is ctypes.  There is NO CPU fallback: if the shared library is missing the
import of any compute entry point raises, and `lbm_create` itself refuses non-sm_100 devices.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LBM_B200_LIB", os.path.join(_HERE, "liblbm_b200.so"))   # env override: perf bisects only
CSRC = os.path.join(_HERE, "csrc")

# ---- constants mirrored from include/lbm_b200.h -------------------------------------------
Q = 19
COMPAT_PHYSICAL, COMPAT_REFERENCE = 0, 1
FEAT_WALLS, FEAT_FORCE, FEAT_PHASE, FEAT_LES, FEAT_POROUS, FEAT_DRIVE, FEAT_STRICT = 1, 2, 4, 8, 16, 32, 64
FLAG_SOLID, FLAG_FILTER, FLAG_LES, FLAG_NEAR = 1, 2, 4, 8


class LbmParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("nz_global", C.c_int), ("z0", C.c_int), ("zghost", C.c_int),
                ("periodic", C.c_int), ("compat", C.c_int), ("features", C.c_int),
                ("tau_water", C.c_float), ("tau_air", C.c_float), ("gravity_lu", C.c_float),
                ("cs_smag", C.c_float), ("tau_min", C.c_float), ("tau_max", C.c_float),
                ("porous_darcy", C.c_float), ("porous_forch", C.c_float),
                ("K_lu", C.c_float), ("beta_lu", C.c_float), ("c_darcy", C.c_float), ("c_forch", C.c_float),
                ("vec", C.c_int), ("block", C.c_int), ("drive_max_force", C.c_float), ("drive_scale", C.c_float),
                ("mrt_magic", C.c_float)]


class LbmFields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("f_src", "f_dst", "rho", "u_src", "u_dst", "body_force", "phase", "blockage", "flags", "rho_src")]


class LbmParticles(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("pos", "vel", "radius", "mass", "active", "drag_new", "drag_old", "drag",
                 "u_fluid", "reynolds", "cd", "cell")] + [("n", C.c_int)]


class LbmPour(C.Structure):
    _fields_ = [("pour_x", C.c_float), ("pour_y", C.c_float), ("radius", C.c_float), ("pour_z", C.c_int),
                ("velocity", C.c_float), ("flow_rate", C.c_float), ("dt", C.c_float)]


class LbmParticleBounds(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("center_x", "center_y", "bottom_z", "bottom_radius_lu", "top_radius_lu",
                                         "cup_height_lu", "max_coordinate", "nz_minus_5")]


# every symbol include/lbm_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "lbm_version": (C.c_int, []),
    "lbm_last_error": (C.c_char_p, [_P]),
    "lbm_create": (C.c_int, [C.POINTER(_P), C.c_int, C.POINTER(LbmParams)]),
    "lbm_set_params": (C.c_int, [_P, C.POINTER(LbmParams)]),
    "lbm_destroy": (None, [_P]),
    "lbm_launch_count": (C.c_longlong, [_P]),
    "lbm_populations_changed": (C.c_int, [_P]),
    "lbm_selftest_math": (C.c_int, [_P, C.POINTER(C.c_ulonglong), _P]),
    "lbm_init_equilibrium": (C.c_int, [_P, _P, _P, _P, C.c_float, C.POINTER(C.c_float), _P]),
    "lbm_build_v60_geometry": (C.c_int, [_P, _P, _P, C.POINTER(C.c_float), _P]),
    "lbm_pack_flags": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "lbm_step": (C.c_int, [_P, C.POINTER(LbmFields), C.c_int, C.c_int, _P, _P]),
    "lbm_macroscopic": (C.c_int, [_P, C.POINTER(LbmFields), _P]),
    "lbm_face_bc": (C.c_int, [_P, C.POINTER(LbmFields), _P]),
    "lbm_export_f": (C.c_int, [_P, _P, _P, _P, _P]),
    "lbm_import_f": (C.c_int, [_P, _P, _P, _P, _P]),
    "lbm_pressure_gradient_force": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_float, _P]),
    "lbm_pressure_gradient_force_set": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_float, _P]),
    "lbm_forchheimer_force": (C.c_int, [_P, _P, _P, _P, C.c_float, _P]),
    "lbm_density_drive": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_float, C.c_float, C.c_float, _P]),
    "lbm_field_statistics": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "lbm_add_reaction_force": (C.c_int, [_P, _P, _P, _P, _P]),
    "lbm_surface_tension": (C.c_int, [_P] * 11 + [C.c_float, _P]),
    "lbm_surface_tension_gradients": (C.c_int, [_P] * 7),
    "lbm_surface_tension_curvature_force": (C.c_int, [_P] * 9 + [C.c_float, _P]),
    "lbm_surface_tension_body_force": (C.c_int, [_P] * 7 + [C.c_float, _P]),
    "lbm_chemical_potential": (C.c_int, [_P, _P, _P, _P, C.c_float, _P]),
    "lbm_apply_surface_tension": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "lbm_phase_field_step": (C.c_int, [_P] * 7 + [C.c_float, C.c_float, C.c_double, C.c_double, _P]),
    "lbm_density_from_phase": (C.c_int, [_P, _P, _P, _P, C.c_double, C.c_double, _P]),
    "lbm_pouring_force": (C.c_int, [_P, C.POINTER(LbmPour), _P, _P, _P]),
    "lbm_pouring_phase_change": (C.c_int, [_P, C.POINTER(LbmPour), _P, _P, _P]),
    "lbm_particles_couple": (C.c_int, [_P, _P, _P, C.POINTER(LbmParticles), C.c_float, C.c_float, C.c_float, _P]),
    "lbm_particles_couple_sparse": (C.c_int, [_P, _P, _P, C.POINTER(LbmParticles), C.c_float, C.c_float, C.c_float, _P]),
    "lbm_particles_couple_slab": (C.c_int, [_P, _P, _P, C.POINTER(LbmParticles), C.c_float, C.c_float, C.c_float, C.c_int, _P]),
    "lbm_particles_under_relax": (C.c_int, [_P, C.POINTER(LbmParticles), C.c_float, _P]),
    "lbm_particles_advance": (C.c_int, [_P, C.POINTER(LbmParticles), _P, C.POINTER(LbmParticleBounds), C.c_float, _P, _P]),
    "lbm_particles_fluid_forces": (C.c_int, [_P, _P, C.POINTER(LbmParticles), _P, C.c_double, C.c_double, C.c_double, _P, _P]),
    "lbm_particles_block_at_filter": (C.c_int, [_P, C.POINTER(LbmParticles), _P, _P, C.c_float, C.c_float, C.c_uint, _P]),
    "lbm_filter_dynamic_resistance": (C.c_int, [_P, _P, _P, _P, _P]),
    "lbm_nccl_unique_id": (C.c_int, [_P]),
    "lbm_attach_nccl": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "lbm_halo_exchange": (C.c_int, [_P, _P, _P, _P]),
    "lbm_halo_exchange_field": (C.c_int, [_P, _P, _P, _P]),
}


def build(force: bool = False, jobs: int | None = None) -> str:
    """Compile csrc/*.cu for sm_100a with the in-tree Makefile (nvcc cross-compiles without a GPU)."""
    jobs = jobs or os.cpu_count() or 4
    cmd = ["make", "-C", CSRC, "-s", f"-j{jobs}"] + (["-B"] if force else [])
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load liblbm_b200.so and attach the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(pour_over_coffee_lbm_b200 has no CPU or PyTorch fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib
